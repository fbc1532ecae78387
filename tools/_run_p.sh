cd $GRAFT_REPO_ROOT
for w in er zinc; do
python tools/sent_only.py $w > gpurun_out/r2p_plain_$w.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"sent_walk_kernel|sent_emit_kernel" -s 2 -c 2 -f -o gpurun_out/r2p_$w python tools/sent_only.py $w > gpurun_out/r2p_ncu_$w.log 2>&1
tail -n 1 gpurun_out/r2p_plain_$w.log gpurun_out/r2p_ncu_$w.log
done
