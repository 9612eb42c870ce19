#!/bin/bash
# Round 2, session 3, GPU call: ncu --set full of the ordering gather kernels (second call: warm).
mkdir -p gpurun_out
export COSMICGPU_LIB=$PWD/cosmic_profiles_b200/libcosmicgpu_dev_q.so
python profiles/ncu_gather.py || exit 1
ncu --set full --clock-control none --import-source on -k 'regex:gather_bucketed|bucket_hist_kernel|tile_tensor_kernel' -s 3 -c 3 -f -o /tmp/r3_gather python profiles/ncu_gather.py > gpurun_out/r3_ncu_gather.log 2>&1
ncu -i /tmp/r3_gather.ncu-rep --page details > gpurun_out/r3_ncu_gather_details.txt 2>/dev/null
ncu -i /tmp/r3_gather.ncu-rep --page source --csv -k regex:gather_bucketed > gpurun_out/r3_ncu_gather_source.csv 2>/dev/null
ls -la gpurun_out | tail -4
