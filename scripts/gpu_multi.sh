# usage: bash scripts/gpu_multi.sh N [tests] [c4] [c5]  -- multi-GPU tests and bench lines at N GPUs (gpurun --gpus N)
N=${1:-2}; shift
mkdir -p gpurun_out
nvidia-smi -L | head -$N; nproc
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
for a in "$@"; do
if [ "$a" = "tests" ]; then
timeout 2400 python -m pytest tests -m gpu_multi -q --timeout 1500 --durations=5 > gpurun_out/pytest_multi_n$N.log 2>&1; echo pytest_multi rc=$?; tail -15 gpurun_out/pytest_multi_n$N.log
fi
if [ "$a" = "c4" ]; then
timeout 900 $RUN --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_c4_n$N.json 2> gpurun_out/bench_c4_n$N.err; echo bench_c4 rc=$?
tail -3 gpurun_out/bench_c4_n$N.err | cut -c1-400
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_c4_n$N.json').read().strip().splitlines()[-1])
    v=d['view_sharding']
    print('C4 N', d['n_gpus'], 'ms/step', round(d['ms_per_step'],4), 'fps', round(d['frames_per_second'],1), 'G/s', round(d['value']/1e9,3), 'e2e fps', round(d['e2e']['frames_per_second'],1), 'verified', v['consumed_frames_verified'], 'timeouts', v['flow_control_timeouts'])
    print(' per-rank render ms/step', v.get('per_rank_render_stream_ms_per_step'), 'consumer', v.get('rank0_consumer_stream_ms_per_step'))
except Exception as e:
    print('parse failed', e)
PY
fi
if [ "$a" = "c4float" ]; then
timeout 900 $RUN --master-port 29514 bench.py --gpus $N --steps 20 --warmup 5 --deliver float > gpurun_out/bench_c4_float_n$N.json 2> gpurun_out/bench_c4_float_n$N.err; echo bench_c4_float rc=$?
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_c4_float_n$N.json').read().strip().splitlines()[-1])
    v=d['view_sharding']
    print('C4 float N', d['n_gpus'], 'ms/step', round(d['ms_per_step'],4), 'fps', round(d['frames_per_second'],1), 'e2e fps', round(d['e2e']['frames_per_second'],1), 'verified', v['consumed_frames_verified'])
    print(' per-rank render ms/step', v.get('per_rank_render_stream_ms_per_step'), 'consumer', v.get('rank0_consumer_stream_ms_per_step'))
except Exception as e:
    print('parse failed', e)
PY
fi
if [ "$a" = "c4ref" ]; then
timeout 900 $RUN --master-port 29513 bench.py --impl reference --gpus $N --steps 5 --warmup 1 > gpurun_out/bench_c4_ref_n$N.json 2> gpurun_out/bench_c4_ref_n$N.err; echo bench_c4_ref rc=$?; cut -c1-300 gpurun_out/bench_c4_ref_n$N.json
fi
if [ "$a" = "c5" ]; then
timeout 1200 $RUN --master-port 29512 bench.py --gpus $N --config C5 --shard rows --steps 10 --warmup 3 > gpurun_out/bench_c5_n$N.json 2> gpurun_out/bench_c5_n$N.err; echo bench_c5 rc=$?
tail -3 gpurun_out/bench_c5_n$N.err | cut -c1-400
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_c5_n$N.json').read().strip().splitlines()[-1])
    t=d['tile_row_sharding']
    print('C5 N', d['n_gpus'], 'ms/frame', round(d['ms_per_step'],4), 'e2e ms', round(d['e2e']['ms_per_step'],4), 'split', t['split'], 'time imbalance', round(t['time_imbalance_max_over_mean'],3), 'verified', t['assembled_frame_equals_single_gpu_frame'], 'timeouts', t['flow_control_timeouts'])
    print(' band ms', t['band_ms_standalone'], [(x['split'], x['max_ms']) for x in t['splits_tried']])
    print(' per-rank ms', t['per_rank_ms_per_frame']); print(' stages rank0', t['per_rank_stage_ms'][0])
except Exception as e:
    print('parse failed', e)
PY
fi
done
