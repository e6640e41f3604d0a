mkdir -p gpurun_out
for v in 0 1 2 3 4 5 6; do LCGS_SORT_VARIANT=$v timeout 300 python scripts/tune_sort.py 2>&1 | tail -1; done | tee gpurun_out/tune_sort.log
