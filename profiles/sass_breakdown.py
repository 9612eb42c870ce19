"""Instruction-level breakdown of one kernel of an ncu report (source page, SASS view): opcode mix, stall reasons, the
share of the fp64 pipe, and the split between the hottest loop (the streaming tile loop) and the rest (the per-iteration
step).  Usage: python profiles/sass_breakdown.py gpurun_out/r2g_morph.ncu-rep morph_warp > profiles/<tag>.txt"""
import collections
import csv
import io
import re
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--kernel-name", "regex:" + kern],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
name = rows[0][1]
hdr = rows[1]
iS, iN, iT, iSm = (hdr.index(k) for k in ("Source", "Instructions Executed", "Thread Instructions Executed", "# Samples"))
data = []
for r in rows[2:]:
    try:
        data.append((r[iS].strip(), int(r[iN]), int(r[iT]), int(r[iSm]), r))
    except (ValueError, IndexError):
        pass
data = data[:len(data) // 2] if len(data) % 2 == 0 and data[0][0] == data[len(data) // 2][0] else data   # the csv lists the kernel twice
tot = sum(d[1] for d in data)
totS = sum(d[3] for d in data)
print(name)
print("SASS lines %d, warp instructions executed %.4g, stall samples %d" % (len(data), tot, totS))
FP64 = re.compile(r"\b(DFMA|DMUL|DADD|DSETP|DMNMX)")
f64 = sum(d[1] for d in data if FP64.search(d[0]))
print("fp64-pipe instructions: %.4g = %.1f %% of all warp instructions" % (f64, 100 * f64 / tot))
byop = collections.Counter()
for s, n, t, sm, _ in data:
    m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_]+)", s)
    byop[m.group(2) if m else s[:8]] += n
print("\nopcode mix (share of warp instructions):")
print("  " + "  ".join("%s %.1f%%" % (op, 100 * n / tot) for op, n in byop.most_common(16)))
cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" in h]
st = collections.Counter()
for *_, r in data:
    for i, h in cols:
        try:
            st[h.replace(" (Not Issued)", "")] += int(r[i])
        except ValueError:
            pass
ss = sum(st.values())
print("\nstall reasons of warps that could not issue (share of not-issued samples):")
print("  " + "  ".join("%s %.1f%%" % (k.replace("stall_", ""), 100 * v / ss) for k, v in st.most_common(8)))
cmax = max(d[1] for d in data)
hot = [d for d in data if d[1] > 0.5 * cmax]
print("\nhottest loop (lines executed more than half as often as the most executed one = the tile loop):")
print("  %d lines, %.1f %% of the instructions, %.1f %% of the samples, %.1f %% of the fp64 instructions; executed %.3g times"
      % (len(hot), 100 * sum(d[1] for d in hot) / tot, 100 * sum(d[3] for d in hot) / totS,
         100 * sum(d[1] for d in hot if FP64.search(d[0])) / f64, cmax))
