#!/bin/bash
# Round 2, GPU call 14: big-object scan/statistics kernels: GPU suite, then configs 4 and the headline
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -x > gpurun_out/r2r_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2r_pytest.log
tail -4 gpurun_out/r2r_pytest.log
for c in 4; do
python bench.py --config $c > gpurun_out/r2r_bench_config$c.json 2> gpurun_out/r2r_bench_config$c.err; echo "config $c rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/r2r_bench_config$c.json")); print($c, d["ms_per_step"], d["value"], d["e2e"]["value"], d["detail"]["kernel_ms"], d["detail"]["gather_ms"])
PY
done
