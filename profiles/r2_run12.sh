#!/bin/bash
# Round 2, GPU call 12: full GPU suite + headline bench with the shell-pair streaming block (default build)
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -x > gpurun_out/r2p_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2p_pytest.log
tail -4 gpurun_out/r2p_pytest.log
python bench.py > gpurun_out/r2p_bench.json 2> gpurun_out/r2p_bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/r2p_bench.json")); print(d["ms_per_step"], d["breakdown_ms_per_step_rank0"], d["roofline"]["frac"], d["e2e"]["ms_per_step"])
PY
