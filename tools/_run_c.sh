cd $GRAFT_REPO_ROOT
export GTOK_ZINC_TMA=1
python tools/probe_zinc.py > gpurun_out/r2c_plain1.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:zinc_tpm_kernel -s 27 -c 2 -f -o gpurun_out/r2c_tma1 python tools/probe_zinc.py > gpurun_out/r2c_ncu1.log 2>&1
tail -3 gpurun_out/r2c_plain1.log; tail -3 gpurun_out/r2c_ncu1.log
