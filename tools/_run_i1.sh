cd $GRAFT_REPO_ROOT
timeout 600 python tools/sweep_large.py 32000000 2>/dev/null | grep ibtt_collate | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print({k:(round(v,3) if isinstance(v,float) else v) for k,v in d.items() if k in ('kind','N','B','ms','frac')})
"
timeout 900 python -m pytest tests -m gpu -x -q -k "ibtt or sweep or guard" 2>&1 | tail -2
