# usage: bash scripts/gpu_multi.sh N   -- multi-GPU correctness check + bench at N GPUs
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L | head -$N
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 scripts/multi_gpu_check.py > gpurun_out/multi_check_n$N.json 2> gpurun_out/multi_check_n$N.err; echo check rc=$?; tail -2 gpurun_out/multi_check_n$N.json; tail -3 gpurun_out/multi_check_n$N.err | cut -c1-300
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo bench rc=$?
tail -3 gpurun_out/bench_n$N.err | cut -c1-300
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_n$N.json').read().strip().splitlines()[-1])
print('N', d['n_gpus'], 'ms/step', round(d['ms_per_step'],3), 'fps', round(d['frames_per_second'],1), 'G/s', round(d['value']/1e9,3), 'e2e G/s', round(d['e2e']['value']/1e9,3))
PY
