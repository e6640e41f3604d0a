# usage: bash scripts/ncu_full.sh <name> <kernel-regex> <skip> <count>
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k "regex:${2:-blend_kernel|onesweep_pass}" -s ${3:-7} -c ${4:-7} -f -o gpurun_out/${1:-prof} \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
echo ncu rc=$?
