cd $GRAFT_REPO_ROOT
python tools/sent_sparse_only.py 256 > gpurun_out/r2y_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"sent_walkl_kernel|sent_emitl_kernel" -s 2 -c 2 -f -o gpurun_out/r2y_walkl_256 python tools/sent_sparse_only.py 256 > gpurun_out/r2y_ncu.log 2>&1
tail -n 1 gpurun_out/r2y_plain.log gpurun_out/r2y_ncu.log
