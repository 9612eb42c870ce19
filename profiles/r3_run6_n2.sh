#!/bin/bash
# Round 2, session 3, 2 GPUs of one box: the multi-device tests and the torchrun bench at N = 2 with the final kernels.
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r3_gpus_n2.txt
(python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -3)
n=2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2961$n bench.py --gpus $n --steps 10 --warmup 3 --no-cpu-baseline --no-dropin > gpurun_out/r3_bench_n$n.json 2> gpurun_out/r3_bench_n$n.err; echo "bench n$n rc=$?"
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r3_bench_n$n.json")); print($n, d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"], d["e2e"]["value"], d.get("multi_gpu",{}).get("step_ms_max_over_mean"))
except Exception as e: print("ERR", e)
PY
