#!/bin/bash
# Round 2, session 3, GPU call 3: kernel 1c variants e (11 pass warps, 1 CTA/SM), f (fp32 pre-classification), g (shell pairs),
# h (prefix lengths in shared memory + fp32 bucket search), i (= e + h), k (kernel 1 with the fp32 bucket search).
mkdir -p gpurun_out
for v in ${VARIANTS:-e f g h i k}; do
  export COSMICGPU_LIB=$PWD/cosmic_profiles_b200/libcosmicgpu_dev_$v.so
  timeout 300 python -m pytest tests/test_gpu_full_size.py -x -q > gpurun_out/r3_full_$v.log 2>&1; echo "variant $v full-size rc=$?"; tail -1 gpurun_out/r3_full_$v.log
  timeout 300 python bench.py --steps 3 --warmup 3 > gpurun_out/r3_bench_$v.json 2> gpurun_out/r3_bench_$v.err; echo "variant $v bench rc=$?"
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r3_bench_$v.json")); print("$v", d["ms_per_step"], d["breakdown_ms_per_step_rank0"]["local_kernel_ms"], d["breakdown_ms_per_step_rank0"]["global_kernel_ms"])
except Exception as e: print("$v ERR", e)
PY
done
