#!/bin/bash
# Round 2, GPU call 10: per-warp cost model (us per tile / per step / per task) of the warp kernel at 12, 8 and 4
# resident warps per SM, and with the split quadratic form.
mkdir -p gpurun_out
for occ in 3 2 1; do
COSMICGPU_LIB=$PWD/variants/lib_trace.so python profiles/task_trace.py r2k_occ$occ warp_ctas_per_sm=$occ 2>&1 | grep -v "^   task" | tee gpurun_out/r2k_occ$occ.log
done
COSMICGPU_LIB=$PWD/variants/lib_trace_split.so python profiles/task_trace.py r2k_split 2>&1 | grep -v "^   task" | tee gpurun_out/r2k_split.log
rm -f gpurun_out/r2k_*_trace.npz
