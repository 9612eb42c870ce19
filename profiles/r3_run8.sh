#!/bin/bash
# Round 2, session 3, last call: launch list and pipe metrics of the final build (coalesced bucket scan included), ncu --set full
# of the ordering-gather kernels.  (Numbers printed under ncu are never bench values.)
mkdir -p gpurun_out
T=r3f
python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"
bash profiles/capture.sh ${T} launches > gpurun_out/${T}_capture.log 2>&1; echo "capture rc=$?"
B="python bench.py --no-cpu-baseline --no-pipeline --steps 1 --warmup 3"
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active \
    --clock-control none -k 'regex:morph_|gather_bucketed|bucket_hist_kernel|bucket_scan_kernel|tile_tensor' -s 27 -c 9 --csv --log-file gpurun_out/${T}_pipe_metrics.csv $B > gpurun_out/${T}_ncu_metrics.log 2>&1
python profiles/pipe_metrics.py gpurun_out/${T}_pipe_metrics.csv ${T}
ncu --set full --clock-control none -k 'regex:gather_bucketed|bucket_hist_kernel' -s 6 -c 2 -f -o /tmp/${T}_gather $B > gpurun_out/${T}_ncu_gather.log 2>&1
ncu -i /tmp/${T}_gather.ncu-rep --page details > gpurun_out/${T}_ncu_gather_details.txt 2>/dev/null
python -c "
import json; d=json.load(open('gpurun_out/${T}_bench.json')); print(d['ms_per_step'], d['value'], d['breakdown_ms_per_step_rank0'], d['e2e']['ms_per_step'], d['roofline']['frac'], d['roofline']['hbm']['physical_frac'])"
