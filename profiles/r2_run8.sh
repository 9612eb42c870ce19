#!/bin/bash
# Round 2, GPU call 8: A/B of streaming-loop variants (variants/lib_<tag>.so built with different -D flags; see
# profiles/README.md).  Short resident-only bench per variant.
mkdir -p gpurun_out
for so in variants/lib_*.so; do
  tag=$(basename $so .so); tag=${tag#lib_}
  COSMICGPU_LIB=$PWD/$so python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-dropin --no-pipeline > gpurun_out/r2i_bench_$tag.json 2> gpurun_out/r2i_bench_$tag.err; echo "$tag rc=$?"
  python - <<PY
import json
d=json.load(open("gpurun_out/r2i_bench_$tag.json")); print("$tag", round(d["ms_per_step"],3), d["breakdown_ms_per_step_rank0"], round(d["roofline"]["frac"],3))
PY
done
