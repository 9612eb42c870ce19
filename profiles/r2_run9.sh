#!/bin/bash
# Round 2, GPU call 9: per-task timeline of the local-shape kernels (diagnostics build with -DCPG_TASK_TRACE)
mkdir -p gpurun_out
tag=${1:-r2j}; shift
COSMICGPU_LIB=$PWD/variants/lib_trace.so python profiles/task_trace.py $tag "$@" 2>&1 | tee gpurun_out/${tag}_trace.log
