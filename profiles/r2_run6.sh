#!/bin/bash
# Round 2, GPU call 6: batched-step warp kernel + single-stepper cluster kernel: full GPU suite, then bench with the
# batched step on (default) and off.
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -x > gpurun_out/r2h_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2h_pytest.log
tail -4 gpurun_out/r2h_pytest.log
for o in batched_step=1 batched_step=0; do
python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-dropin --no-pipeline --opt $o > gpurun_out/r2h_bench_$o.json 2>gpurun_out/r2h_bench_$o.err; echo "rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/r2h_bench_$o.json")); print("$o", d["ms_per_step"], d["breakdown_ms_per_step_rank0"], d["roofline"]["frac"])
PY
done
