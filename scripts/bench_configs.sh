mkdir -p gpurun_out
for c in C1 C2 C5; do
  extra=""; if [ $c = C5 ]; then extra="--no-cpu-baseline --steps 5 --warmup 3"; else extra="--steps 20 --warmup 5"; fi
  timeout 900 python bench.py --config $c $extra > gpurun_out/bench_$c.json 2> gpurun_out/bench_$c.err; echo $c rc=$?
  python - <<PY
import json
d=json.loads(open('gpurun_out/bench_$c.json').read().strip().splitlines()[-1])
print('$c ms/frame', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), 'N', d['num_rendered'], {k:v['ms'] for k,v in d['stages'].items()}, 'cpu', d.get('cpu_baseline',{}).get('ms_per_frame'))
PY
  tail -2 gpurun_out/bench_$c.err | cut -c1-300
done
