#!/usr/bin/env python3
"""Per-kernel sums of ONE Groth16 proof (or one MSM step) out of an ncu launch list
(`ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file X ...`).
usage: summarize_launches.py X.csv groth16|msm  > summary.csv"""
import csv, sys, collections, re

rows = list(csv.reader(open(sys.argv[1], errors="ignore")))
hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
h = rows[hi]
ki, vi = h.index("Kernel Name"), h.index("Metric Value")
launches = []
for r in rows[hi + 1:]:
    if len(r) <= vi:
        continue
    try:
        ns = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    name = re.sub(r"^void ", "", r[ki])
    name = re.sub(r"^ozl::", "", name).split("<")[0].split("(")[0]
    launches.append((name, ns / 1e6))
mode = sys.argv[2] if len(sys.argv) > 2 else "groth16"
first = "k_spmv" if mode == "groth16" else "k_count"
idx = [i for i, (n, _) in enumerate(launches) if n == first and (i == 0 or launches[i - 1][0] != first)]
start = idx[-1] if idx else 0
end = len(launches)
if mode == "groth16":   # three consecutive k_spmv open a proof; take the 4th one (after the warm-up, single prover)
    idx = [i for i in range(len(launches) - 2) if all(launches[i + j][0] == "k_spmv" for j in range(3))]
    pick = min(3, len(idx) - 1)
    start = idx[pick]
    end = idx[pick + 1] if pick + 1 < len(idx) else len(launches)
if mode == "msm":   # one device-resident step: k_count .. next k_count, with one accumulation and one bucket reduction
    bounds = [i for i, (n, _) in enumerate(launches) if n == "k_count"] + [len(launches)]
    segs = [(a, b) for a, b in zip(bounds, bounds[1:])]
    def ok(a, b):
        names = [n for n, _ in launches[a:b]]
        return names.count("k_accumulate_tma") == 1 and names.count("k_bucket_reduce") == 1 and "k_bench_mul" not in names
    cands = [sg for sg in segs if ok(*sg)]
    # the timed configuration is the one measured last; its device-resident steps have the longest accumulation
    # (the batches of a host-scalar call accumulate a part of the points each)
    def nsc(sg): return [n for n, _ in launches[sg[0]:sg[1]]].count("k_scatter_window")
    def tacc(sg): return max(v for n, v in launches[sg[0]:sg[1]] if n == "k_accumulate_tma")
    cands = [sg for sg in cands if nsc(sg) == nsc(cands[-1])]
    start, end = max(cands, key=tacc)
    # the step ends with k_final
    for i in range(start, end):
        if launches[i][0] == "k_final":
            end = i + 1
            break
sel = launches[start:end]
tot = sum(v for _, v in sel)
agg, cnt = collections.OrderedDict(), collections.Counter()
for n, v in sel:
    agg[n] = agg.get(n, 0.0) + v
    cnt[n] += 1
print("kernel,launches,ms,share_pct")
for n, v in sorted(agg.items(), key=lambda x: -x[1]):
    print(f"{n},{cnt[n]},{v:.3f},{100 * v / tot:.2f}")
print(f"TOTAL,{len(sel)},{tot:.3f},100")
