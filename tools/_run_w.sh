cd $GRAFT_REPO_ROOT
echo "=== SENT tests"; timeout 2400 python -m pytest tests -m gpu -x -q -k "sent or shim or agtt or pin or smoke or sweep or guard" 2>&1 | tail -12
echo "=== sweep"; timeout 900 python tools/sweep_large.py 32000000 > gpurun_out/r2w_sweep.jsonl 2>gpurun_out/r2w_sweep.err; grep "sent" gpurun_out/r2w_sweep.jsonl | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print({k:(round(v,3) if isinstance(v,float) else v) for k,v in d.items() if not isinstance(v,(dict,list))})
"
