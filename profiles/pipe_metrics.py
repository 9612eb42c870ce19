"""Parse the CSV of `ncu --metrics ... --csv --log-file` (one row per kernel launch and metric) into
{kernel: {metric: [value, unit]}} (first launch of every kernel name) and derive the traffic / fp64-pipe records
bench.py reads (profiles/<tag>_traffic.json, <tag>_fp64_pipe.json)."""
import csv
import json
import sys

src, tag = sys.argv[1], sys.argv[2]
rows = [r for r in csv.reader(open(src, errors="replace")) if len(r) > 5]
hdr = next(r for r in rows if "Kernel Name" in r)
ik, im, iu, iv, iid = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Unit"), hdr.index("Metric Value"), hdr.index("ID")
out, first_id = {}, {}
for r in rows:
    if r is hdr or not r[iid].isdigit():
        continue
    name = r[ik]
    if name not in first_id:
        first_id[name] = r[iid]
    if first_id[name] != r[iid]:
        continue
    out.setdefault(name, {})[r[im]] = [r[iv].replace(",", ""), r[iu]]
json.dump(out, open("gpurun_out/%s_pipe_metrics.json" % tag, "w"), indent=1)


def pick(pattern):
    for k, v in out.items():
        if pattern in k:
            return k, v
    return None, None


def nbytes(v):
    return float(v["dram__bytes_read.sum"][0]) + float(v["dram__bytes_write.sum"][0])


kw, vw = pick("morph_wduo_kernel<4")
if vw is None:
    kw, vw = pick("morph_warp_kernel<4")
kc, vc = pick("morph_cluster_kernel<4")
if vw and vc:
    json.dump({"workload_halos": 10000, "dram_bytes_per_local_shape_launch_set": nbytes(vw) + nbytes(vc),
               "per_kernel": {kw: nbytes(vw), kc: nbytes(vc)},
               "source": "profiles/%s_pipe_metrics.json (ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none)" % tag,
               "note": "dram__bytes_read.sum + dram__bytes_write.sum of the two local-shape kernels of one step"},
              open("gpurun_out/%s_traffic.json" % tag, "w"), indent=1)
    m = "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"
    json.dump({"warp_kernel_pct": float(vw[m][0]), "cluster_kernel_pct": float(vc[m][0]), "dfma_ceiling_pct": 92.0,
               "warp_kernel": kw, "cluster_kernel": kc,
               "source": "%s under ncu: the two local-shape kernels on configs[1] (profiles/%s_pipe_metrics.json); the pure-DFMA "
                         "microbenchmark profiles/micro/fp64_lanes.cu reads 92.0 %% on the same counter (profiles/r2_fp64_lanes_ncu.txt)" % (m, tag)},
              open("gpurun_out/%s_fp64_pipe.json" % tag, "w"), indent=1)
for k, v in out.items():
    print(k[:60], {a: b[0] for a, b in v.items()})
