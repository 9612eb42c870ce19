#!/bin/bash
# Round 2, GPU call 3 (1 GPU): full GPU suite on the default (all-fp64) build, the bench line, launch list and one
# ncu --set full capture of the two local-shape morphology kernels.
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2c_pytest.log
tail -3 gpurun_out/r2c_pytest.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err; echo "bench rc=$?"
B="python bench.py --no-cpu-baseline --no-pipeline --no-dropin --steps 1 --warmup 3"
$B > gpurun_out/r2c_bench_plain.json 2> gpurun_out/r2c_bench_plain.err || exit 1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2c_launches.csv $B > gpurun_out/r2c_ncu_launch.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k 'regex:morph_(warp|cluster)_kernel' -s 12 -c 2 -f -o /tmp/r2c_morph $B > gpurun_out/r2c_ncu_full.log 2>&1
ncu -i /tmp/r2c_morph.ncu-rep --page raw --csv > gpurun_out/r2c_morph_raw.csv 2>/dev/null
ncu -i /tmp/r2c_morph.ncu-rep --page source --csv --kernel-name regex:morph_warp > gpurun_out/r2c_src_warp.csv 2>/dev/null
ls -la gpurun_out | tail -12
