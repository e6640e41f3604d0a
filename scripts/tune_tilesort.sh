# usage: VARIANTS="2 3 4" bash scripts/tune_tilesort.sh   -- tile sort timing per onesweep geometry (+ phase timers)
mkdir -p gpurun_out
for v in ${VARIANTS:-2 3 4}; do
  LCGS_SORT_VARIANT=$v timeout 300 python scripts/tune_tilesort.py 2>&1 | grep -E "variant" | tail -1
done | tee gpurun_out/tune_tilesort.log
LCGS_SORT_DEBUG=1 LCGS_SORT_DEBUG_PRINT=1 timeout 300 python scripts/tune_tilesort.py 2>&1 | grep -E "sort dbg" | tail -1 | tee -a gpurun_out/tune_tilesort.log
