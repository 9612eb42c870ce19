#!/bin/bash
# Round 2, GPU call 2: full GPU suite on the default build + A/B of kernel variants (COSMICGPU_LIB selects the build).
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2b_pytest.log
tail -15 gpurun_out/r2b_pytest.log
for v in A B C D E F; do
  lib=$PWD/cosmic_profiles_b200/libcosmicgpu_$v.so
  [ $v = A ] && lib=$PWD/cosmic_profiles_b200/libcosmicgpu.so
  [ -f $lib ] || continue
  COSMICGPU_LIB=$lib timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-dropin --no-pipeline > gpurun_out/r2b_bench_$v.json 2> gpurun_out/r2b_bench_$v.err; echo "bench $v rc=$?"
  python - <<PY
import json
d=json.load(open("gpurun_out/r2b_bench_$v.json"))
print("$v", "step %.2f ms" % d["ms_per_step"], d["breakdown_ms_per_step_rank0"])
PY
done
