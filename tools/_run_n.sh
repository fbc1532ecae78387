cd $GRAFT_REPO_ROOT
python tools/sent_only.py > gpurun_out/r2n_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"sent_walk_kernel|sent_emit_kernel" -s 2 -c 2 -f -o gpurun_out/r2n_walk_er python tools/sent_only.py > gpurun_out/r2n_ncu.log 2>&1
tail -n 2 gpurun_out/r2n_plain.log gpurun_out/r2n_ncu.log
