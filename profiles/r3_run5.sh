#!/bin/bash
# Round 2, session 3, final records on one GPU: headline bench, reference arm, the four other configurations, launch list,
# ncu --set full capture of the local-shape kernels, pipe / traffic metrics of the morphology and gather kernels.
# (Numbers printed under ncu are never bench values.)
mkdir -p gpurun_out
T=r3
(python -m pytest tests -m gpu -x -q 2>&1 | tail -2) ; echo "pytest done"
python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_bench_reference.json 2> gpurun_out/${T}_bench_reference.err; echo "reference rc=$?"
for c in 0 2 3 4; do
python bench.py --config $c > gpurun_out/${T}_bench_config$c.json 2> gpurun_out/${T}_bench_config$c.err; echo "config $c rc=$?"
done
bash profiles/capture.sh ${T} both > gpurun_out/${T}_capture.log 2>&1; echo "capture rc=$?"
ncu -i gpurun_out/${T}_morph.ncu-rep --page details > gpurun_out/${T}_ncu_morph_kernels.txt 2>/dev/null
python profiles/sass_breakdown.py gpurun_out/${T}_morph.ncu-rep morph_wduo > gpurun_out/${T}_sass_breakdown_wduo_kernel.txt
python profiles/sass_breakdown.py gpurun_out/${T}_morph.ncu-rep morph_cluster > gpurun_out/${T}_sass_breakdown_cluster_kernel.txt
rm -f gpurun_out/${T}_morph.ncu-rep
B="python bench.py --no-cpu-baseline --no-pipeline --steps 1 --warmup 3"
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active \
    --clock-control none -k 'regex:morph_|gather_bucketed|bucket_hist_kernel|tile_tensor_kernel' -s 21 -c 7 --csv --log-file gpurun_out/${T}_pipe_metrics.csv $B > gpurun_out/${T}_ncu_metrics.log 2>&1
python profiles/pipe_metrics.py gpurun_out/${T}_pipe_metrics.csv ${T}
python - <<PY
import json
for f in ("bench","bench_reference","bench_config0","bench_config2","bench_config3","bench_config4"):
    try:
        d=json.load(open("gpurun_out/${T}_%s.json"%f)); print(f, d.get("ms_per_step"), d.get("value"), (d.get("e2e") or {}).get("value"))
    except Exception as e: print(f, "ERR", e)
PY
