# usage: bash tools/sweep_strips.sh  -- bench the emitter for a few strip geometries (developer aid)
for cfg in "512 8" "1024 8" "512 4" "256 8" "512 16"; do set -- $cfg; echo "== rows $1 cols $2"; NNSDP_STRIP_ROWS=$1 NNSDP_STRIP_COLS=$2 python bench.py --no-cpu --queries 256 --steps 3 --e2e-queries 1 | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l)
        print(d['ms_per_step'], [ (k['kernel'][5:9], round(k['avg_launch_ms'],3), round(k['achieved'])) for k in d['roofline']['emitter_pass']['kernels']])
"; done
