# usage: bash scripts/gpu_round.sh <tag> [ncu]   -- GPU tests + bench (+ optional ncu launch list)
TAG=${1:-x}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/pytest_$TAG.log 2>&1; echo pytest rc=$?; tail -12 gpurun_out/pytest_$TAG.log
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo bench rc=$?
python - <<PY
import json
d=json.load(open('gpurun_out/bench_$TAG.json'))
print('ms/frame', round(d['ms_per_step'],3), 'e2e ms', round(d['e2e']['ms_per_step'],3), 'G/s', round(d['value']/1e9,3))
print({k:v['ms'] for k,v in d['stages'].items()}, d['sort_breakdown'])
print('roofline', d['roofline']['achieved'], d['roofline']['frac'], 'clocks', d['clocks'])
print('cpu', d.get('cpu_baseline',{}).get('ms_per_frame'))
PY
tail -3 gpurun_out/bench_$TAG.err
if [ "$2" = "ncu" ]; then
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench_$TAG.log 2>&1; echo ncu rc=$?
fi
