cd $GRAFT_REPO_ROOT
python tools/sent_only.py > gpurun_out/r2l_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:sent_walk_kernel -s 1 -c 1 -f -o gpurun_out/r2l_walk_er python tools/sent_only.py > gpurun_out/r2l_ncu.log 2>&1
tail -2 gpurun_out/r2l_plain.log gpurun_out/r2l_ncu.log
