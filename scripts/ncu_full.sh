# one ncu --set full capture of one frame's kernels (second frame: skip the first 12 of our launches)
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k "regex:blend_kernel|onesweep_pass|preprocess_fused|duplicate_keys|radix_histogram|scan_inclusive|tile_ranges" -s 12 -c 12 -f -o gpurun_out/${1:-prof_r01} \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
echo ncu rc=$?
ls -la gpurun_out/
