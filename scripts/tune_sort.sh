mkdir -p gpurun_out
for v in ${VARIANTS:-0 1}; do LCGS_SORT_DEBUG=1 LCGS_SORT_DEBUG_PRINT=1 LCGS_SORT_VARIANT=$v timeout 300 python scripts/tune_sort.py 2>&1 | grep -E "variant|sort dbg" | tail -3; done | tee gpurun_out/tune_sort.log
