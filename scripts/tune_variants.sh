# usage: ENVNAME=LCGS_SORT_VARIANT VARIANTS="0 8 9" bash scripts/tune_variants.sh
mkdir -p gpurun_out
for v in $VARIANTS; do echo "$ENVNAME=$v"; env $ENVNAME=$v python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(' ms/frame', round(d['ms_per_step'],3), {k:v['ms'] for k,v in d['stages'].items()}, d['sort_breakdown'])
"; done | tee gpurun_out/tune_variants.log
