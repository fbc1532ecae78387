cd $GRAFT_REPO_ROOT
echo "=== new tests"; timeout 1500 python -m pytest tests -m gpu -x -q -k "append_eos or workers or generator_families or smoke" 2>&1 | tail -5
echo "=== plain"; timeout 600 python tools/sanitize_small.py 2>&1 | tail -2
for tool in memcheck racecheck synccheck; do
  echo "=== compute-sanitizer $tool"
  timeout 1500 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_small.py > gpurun_out/r2u_$tool.log 2>&1; echo rc=$?
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize_small: ok|Error|hazard" gpurun_out/r2u_$tool.log | head -12
done
