cd $GRAFT_REPO_ROOT
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras"
$B > gpurun_out/r2_plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv $B > gpurun_out/r2_ncu_launch.log 2>&1
echo "launch list rc=$?"
$B > gpurun_out/r2_plain_bench2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:zinc_tpm_kernel -s 4 -c 1 -f -o gpurun_out/r2_zinc_collate $B > gpurun_out/r2_ncu_zinc.log 2>&1
echo "zinc full rc=$?"
for w in er zinc; do
python tools/sent_only.py $w > gpurun_out/r2_plain_$w.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"sent_walk_kernel|sent_emit_kernel" -s 2 -c 2 -f -o gpurun_out/r2_sent_$w python tools/sent_only.py $w > gpurun_out/r2_ncu_$w.log 2>&1
echo "sent $w rc=$?"
done
python tools/sent_only.py er fused > gpurun_out/r2_plain_erf.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"sent_emit_kernel" -s 1 -c 1 -f -o gpurun_out/r2_sent_er_fused python tools/sent_only.py er fused > gpurun_out/r2_ncu_erf.log 2>&1
echo "sent er fused rc=$?"
python tools/sent_sparse_only.py 256 > gpurun_out/r2_plain_s256.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"sent_walkm_kernel|sent_emitl_kernel" -s 2 -c 2 -f -o gpurun_out/r2_sent_256 python tools/sent_sparse_only.py 256 > gpurun_out/r2_ncu_s256.log 2>&1
echo "sent 256 rc=$?"
