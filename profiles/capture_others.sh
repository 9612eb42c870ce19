#!/bin/bash
# gpurun -- bash profiles/capture_others.sh <tag>: ncu --set full of one launch of every non-morphology kernel family;
# only the raw-metric CSV comes back (the .ncu-rep of ~50 kernels exceeds the 64 MiB return limit).
tag=${1:-r1b}
python profiles/ncu_others.py --once > gpurun_out/${tag}_others_plain.log 2>&1 || { tail -5 gpurun_out/${tag}_others_plain.log; exit 1; }
ncu --set full --clock-control none \
    -k 'regex:bucket_hist|gather_bucketed|stats_chunk|tile_tensor_kernel|center|dens_|objcat_fill|scan_apply|fit_profiles' \
    -c 40 -f -o /tmp/${tag}_others python profiles/ncu_others.py --once > gpurun_out/${tag}_ncu_others.log 2>&1
ncu -i /tmp/${tag}_others.ncu-rep --page raw --csv > gpurun_out/${tag}_others_raw.csv 2>/dev/null
tail -3 gpurun_out/${tag}_ncu_others.log
ls -la gpurun_out | tail -5
