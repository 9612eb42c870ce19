#!/bin/bash
# Round 2, GPU call 13: shell-variant inner skip: GPU suite, then configs 0 and 4 (the shell-based ones)
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -x > gpurun_out/r2q_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2q_pytest.log
tail -4 gpurun_out/r2q_pytest.log
for c in 0 4; do
python bench.py --config $c > gpurun_out/r2q_bench_config$c.json 2> gpurun_out/r2q_bench_config$c.err; echo "config $c rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/r2q_bench_config$c.json")); print($c, d["ms_per_step"], d["value"], d["e2e"]["value"])
PY
done
