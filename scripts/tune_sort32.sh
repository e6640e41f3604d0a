mkdir -p gpurun_out
for v in 0 1 2 3 4; do echo "sort32 variant $v"; LCGS_SORT32_VARIANT=$v python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(' ms/frame', round(d['ms_per_step'],3), {k:v['ms'] for k,v in d['stages'].items()})
"; done | tee gpurun_out/tune_sort32.log
