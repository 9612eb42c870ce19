#!/bin/bash
# Run on the GPU box (gpurun -- bash profiles/capture.sh <tag> [launches|full|both]): launch list and/or one
# ncu --set full capture of the local-shape morphology kernels of bench.py's workload.
# Numbers printed under ncu are never bench values.
tag=${1:-r1}
what=${2:-both}
set -x
B="python bench.py --no-cpu-baseline --no-pipeline --steps 1 --warmup 3"
$B > gpurun_out/${tag}_bench_plain.json 2> gpurun_out/${tag}_bench_plain.err || exit 1
if [ "$what" != full ]; then
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${tag}_launches.csv $B > gpurun_out/${tag}_ncu_launch.log 2>&1
fi
if [ "$what" != launches ]; then
# a step launches cluster<4>, warp<4> (local shapes) then cluster<1>, warp<1> (global shape); -s 12 = 4th step
ncu --set full --clock-control none --import-source on -k 'regex:morph_(warp|wduo|wasync|cluster)_kernel' -s 12 -c 2 -f -o gpurun_out/${tag}_morph $B > gpurun_out/${tag}_ncu_full.log 2>&1
fi
ls -la gpurun_out
