set -x
python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/pytest2.log 2>&1; echo pytest rc=$?; tail -15 gpurun_out/pytest2.log
python bench.py --steps 20 --warmup 5 > gpurun_out/bench2.json 2> gpurun_out/bench2.err; echo bench rc=$?; cat gpurun_out/bench2.json; tail -5 gpurun_out/bench2.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo ncu rc=$?
