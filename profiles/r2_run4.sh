#!/bin/bash
# Round 2, GPU call 4 (1 GPU): GPU suite, headline bench line, and the other BASELINE configurations (--config 0/2/3/4).
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/r2d_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2d_pytest.log
tail -3 gpurun_out/r2d_pytest.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err; echo "bench rc=$?"
for c in 0 2 3 4; do
  timeout 1200 python bench.py --config $c --steps 5 --warmup 3 > gpurun_out/r2d_bench_config$c.json 2> gpurun_out/r2d_bench_config$c.err; echo "config $c rc=$?"
  tail -2 gpurun_out/r2d_bench_config$c.err
done
timeout 900 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/r2d_bench_reference.json 2> gpurun_out/r2d_bench_reference.err; echo "reference rc=$?"
ls -la gpurun_out | tail -14
