#!/bin/bash
# Round 2, GPU call 1 (gpurun -- bash profiles/r2_run1.sh): GPU test suite, bench line, A/B of the fp32
# pre-classification against the all-fp64 membership test, launch list + ncu full capture of the morphology kernels.
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2a_gpus.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2a_pytest.log
tail -5 gpurun_out/r2a_pytest.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/r2a_bench.json
if [ -f cosmic_profiles_b200/libcosmicgpu_f64.so ]; then
COSMICGPU_LIB=$PWD/cosmic_profiles_b200/libcosmicgpu_f64.so timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-dropin --no-pipeline > gpurun_out/r2a_bench_f64test.json 2> gpurun_out/r2a_bench_f64test.err; echo "bench f64 rc=$?"
fi
B="python bench.py --no-cpu-baseline --no-pipeline --no-dropin --steps 1 --warmup 3"
$B > gpurun_out/r2a_bench_plain.json 2> gpurun_out/r2a_bench_plain.err || exit 1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2a_launches.csv $B > gpurun_out/r2a_ncu_launch.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k 'regex:morph_(warp|cluster)_kernel' -s 12 -c 2 -f -o gpurun_out/r2a_morph $B > gpurun_out/r2a_ncu_full.log 2>&1
ls -la gpurun_out
