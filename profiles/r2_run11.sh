#!/bin/bash
# Round 2, GPU call 11: per-warp cost model of every variants/lib_trace_*.so
mkdir -p gpurun_out
for so in variants/lib_trace_*.so; do
tag=$(basename $so .so); tag=${tag#lib_trace_}
echo "== $tag"
COSMICGPU_LIB=$PWD/$so python profiles/task_trace.py r2l_$tag "$@" 2>&1 | grep -v "^   task" | tee gpurun_out/r2l_$tag.log
done
rm -f gpurun_out/r2l_*_trace.npz
