#!/bin/bash
# Round 2, session 3, GPU call 2: cycle accounting of kernel 1c (variant a traced) + variant d (4 pass warps, 3-stage slots).
mkdir -p gpurun_out
COSMICGPU_LIB=$PWD/cosmic_profiles_b200/libcosmicgpu_dev_t.so timeout 300 python profiles/wasync_trace.py > gpurun_out/r3_trace_a.txt 2>&1; echo "trace rc=$?"; cat gpurun_out/r3_trace_a.txt
for v in d; do
  export COSMICGPU_LIB=$PWD/cosmic_profiles_b200/libcosmicgpu_dev_$v.so
  timeout 300 python bench.py --steps 3 --warmup 3 > gpurun_out/r3_bench_$v.json 2> gpurun_out/r3_bench_$v.err; echo "variant $v bench rc=$?"
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r3_bench_$v.json")); print("$v", d["ms_per_step"], d["breakdown_ms_per_step_rank0"])
except Exception as e: print("$v ERR", e)
PY
done
