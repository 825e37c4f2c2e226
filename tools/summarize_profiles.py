#!/usr/bin/env python
"""Turn the raw ncu outputs brought back in gpurun_out/ into the committed summaries under profiles/:
launch list summary, key metrics of the full-set capture, stall summary of the dominant kernel, traffic.json.

    python tools/summarize_profiles.py [tag]  (build container; needs ncu to read the .ncu-rep)

tag r1  : the dense-ALS kernel set (gpurun_out/prof_r1_final.ncu-rep, launches_r1.csv)
tag r1b : the compact-page kernel set (gpurun_out/prof_r1b.ncu-rep, launches_r1b.csv)  [default]
"""
import collections
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
GO, PR = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
TAG = sys.argv[1] if len(sys.argv) > 1 else "r1b"
REP, LAUNCHES, DOMINANT, STALLS = {
    "r1": ("prof_r1_final.ncu-rep", "launches_r1.csv", "als_kernel<0>", "r1_als_kernel_stalls.txt"),
    "r1b": ("prof_r1b.ncu-rep", "launches_r1b.csv", "als_sparse_kernel", "r1b_als_sparse_kernel_stalls.txt"),
}[TAG]
rep = os.path.join(GO, REP)
for page, dst in (("raw", "raw_final.csv"), ("source", "src_final.csv")):
    with open(os.path.join(GO, dst), "w") as f:
        subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], stdout=f, stderr=subprocess.DEVNULL, check=False)
shutil.copy(os.path.join(GO, LAUNCHES), os.path.join(PR, f"{TAG}_launches_bench.csv"))

rows = [r for r in csv.reader(open(os.path.join(GO, LAUNCHES))) if r and not r[0].startswith("==")]
hdr = rows[0]
idx = {h: i for i, h in enumerate(hdr)}
agg = collections.defaultdict(list)
for r in rows[1:]:
    try:
        v, u = float(r[idx["Metric Value"]]), r[idx["Metric Unit"]]
    except Exception:
        continue
    agg[r[idx["Kernel Name"]].split("(")[0]].append(v / 1e3 if u == "ns" else v)
tot = sum(sum(v) for v in agg.values())
with open(os.path.join(PR, f"{TAG}_launches_summary.csv"), "w") as f:
    f.write("kernel,launches,avg_us,share_pct\n")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        f.write(f"\"{k}\",{len(v)},{sum(v) / len(v):.2f},{100 * sum(v) / tot:.1f}\n")
print(open(os.path.join(PR, f"{TAG}_launches_summary.csv")).read())

rows = list(csv.reader(open(os.path.join(GO, "raw_final.csv"))))
hdr, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit", "launch__shared_mem_per_block", "sm__warps_active.avg.pct", "smsp__issue_active.avg.pct",
        "sm__pipe_fma_cycles_active.avg.pct", "sm__pipe_fmaheavy_cycles_active.avg.pct", "sm__pipe_fmalite_cycles_active.avg.pct",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct", "gpu__dram_throughput.avg.pct", "sm__cycles_elapsed.max",
        "sm__inst_executed_pipe_lsu.avg.pct", "smsp__inst_executed.sum", "sm__throughput.avg.pct",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__warp_issue_stalled"]
keep = [h for h in hdr if any(t in h for t in want) and "peak_sustained" not in h.split(".")[-1]]
ki = hdr.index("Kernel Name")
with open(os.path.join(PR, f"{TAG}_ncu_full_key_metrics.csv"), "w") as f:
    w = csv.writer(f)
    w.writerow(["metric", "unit"] + [r[ki] for r in rows[2:]])
    for h in keep:
        i = hdr.index(h)
        w.writerow([h, units[i]] + [r[i] for r in rows[2:]])


def val(r, k):
    return float(r[hdr.index(k)]) * {"Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "Gbyte": 1e9}[units[hdr.index(k)]]


def dram(name):
    r = [r for r in rows[2:] if name in r[ki]]
    return (val(r[0], "dram__bytes_read.sum"), val(r[0], "dram__bytes_write.sum")) if r else (None, None)


tpath = os.path.join(PR, "traffic.json")
traffic = json.load(open(tpath)) if os.path.exists(tpath) else {}
if TAG == "r1":
    rd, wr = dram("als_kernel<0>")
    traffic.update({"als_kernel_iterate_dram_bytes_per_launch": rd + wr, "dram_bytes_read": rd, "dram_bytes_write": wr})
else:
    for key, name in (("als_sparse_kernel", "als_sparse_kernel"), ("als_sparsify_raw_kernel", "als_sparsify_raw_kernel"),
                      ("fuse_tail_kernel", "fuse_tail_kernel")):
        rd, wr = dram(name)
        if rd is not None:
            traffic.update({f"{key}_dram_bytes_per_launch": rd + wr, f"{key}_dram_bytes_read": rd, f"{key}_dram_bytes_write": wr})
    traffic["source_r1b"] = ("profiles/r1b_ncu_full_key_metrics.csv (ncu --set full, one step, batch 16, scales 8/16/32, raw-matrix inputs); "
                             "writes that were still in the 126 MB L2 when a launch ended reach HBM later as write-backs")
json.dump(traffic, open(tpath, "w"), indent=1)
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print(d["Kernel Name"][:32], {k: d[k] for k in ("gpu__time_duration.sum", "dram__bytes_read.sum", "launch__registers_per_thread",
                                                     "smsp__issue_active.avg.pct_of_peak_sustained_active",
                                                     "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
                                                     "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active",
                                                     "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed") if k in d})

rows = list(csv.reader(open(os.path.join(GO, "src_final.csv"))))
secs = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
st = [i for i in secs if DOMINANT in rows[i][1]][0]
hdr = rows[st + 1]
body = rows[st + 2:min([i for i in secs if i > st] + [len(rows)])]
idx = {h: i for i, h in enumerate(hdr)}
stall = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
ns = sum(int(r[idx["# Samples"]]) for r in body if r[idx["# Samples"]].isdigit())
out = [f"kernel: {rows[st][1]}", f"warp-state samples: {ns}, SASS instructions: {len(body)}"]
totals = {h: sum(int(r[idx[h]]) for r in body if r[idx[h]].isdigit()) for h in stall}
for h, v in sorted(totals.items(), key=lambda kv: -kv[1])[:10]:
    out.append(f"  {h:28s} {v:6d} {100 * v / ns:5.1f}%")
mx = max(int(r[idx["Instructions Executed"]]) for r in body if r[idx["Instructions Executed"]].isdigit())
loop = [r for r in body if r[idx["Instructions Executed"]].isdigit() and int(r[idx["Instructions Executed"]]) >= 0.9 * mx]
out.append(f"main iteration loop: {len(loop)} instructions, {sum(int(r[idx['# Samples']]) for r in loop)} samples; hottest instructions:")
for r in sorted(loop, key=lambda r: -int(r[idx["# Samples"]]))[:25]:
    out.append(f"  {int(r[idx['# Samples']]):4d}  {r[idx['Source']].strip()[:84]:84s} "
               + str({h[6:]: int(r[idx[h]]) for h in stall if r[idx[h]] not in ("0", "") and int(r[idx[h]]) > 5}))
open(os.path.join(PR, STALLS), "w").write("\n".join(out) + "\n")
print("\n".join(out[:14]))
