#!/bin/bash
# Round 2, GPU call 15 (final records, one GPU): headline bench, reference arm, the four other configurations,
# launch list and ncu --set full capture of the local-shape kernels (numbers printed under ncu are never bench values).
mkdir -p gpurun_out
python bench.py > gpurun_out/r2w_bench.json 2> gpurun_out/r2w_bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2w_bench_reference.json 2> gpurun_out/r2w_bench_reference.err; echo "reference rc=$?"
for c in 0 2 3 4; do
python bench.py --config $c > gpurun_out/r2w_bench_config$c.json 2> gpurun_out/r2w_bench_config$c.err; echo "config $c rc=$?"
done
bash profiles/capture.sh r2w both > gpurun_out/r2w_capture.log 2>&1; echo "capture rc=$?"
ncu -i gpurun_out/r2w_morph.ncu-rep --page details > gpurun_out/r2w_ncu_morph_kernels.txt 2>/dev/null
python profiles/sass_breakdown.py gpurun_out/r2w_morph.ncu-rep morph_warp > gpurun_out/r2w_sass_breakdown_warp_kernel.txt
python profiles/sass_breakdown.py gpurun_out/r2w_morph.ncu-rep morph_cluster > gpurun_out/r2w_sass_breakdown_cluster_kernel.txt
rm -f gpurun_out/r2w_morph.ncu-rep
python - <<PY
import json
for f in ("bench","bench_reference","bench_config0","bench_config2","bench_config3","bench_config4"):
    try:
        d=json.load(open("gpurun_out/r2w_%s.json"%f)); print(f, d.get("ms_per_step"), d.get("value"), (d.get("e2e") or {}).get("value"))
    except Exception as e: print(f, "ERR", e)
PY
