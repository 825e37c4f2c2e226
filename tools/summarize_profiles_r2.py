#!/usr/bin/env python
"""Round-2 profile summaries: turn the raw ncu outputs brought back in gpurun_out/ into the committed files under
profiles/ (build container; needs ncu to read the .ncu-rep files).

inputs : gpurun_out/prof_r2_batch16.ncu-rep   ncu --set full, one batch-16 call (tools/profile_als.py)
         gpurun_out/prof_r2_chipfull.ncu-rep  ncu --set full, 29 batches of 16 per launch (tools/profile_big.py 29)
         gpurun_out/launches_r2.csv           ncu --metrics gpu__time_duration.sum of the bench command
outputs: profiles/r2_launches_bench.csv, r2_launches_summary.csv, r2_ncu_full_key_metrics.csv,
         r2_chip_full_key_metrics.csv, r2_als_pages_kernel_stalls.txt, r2_sass_excerpts.txt, traffic.json
"""
import collections
import csv
import json
import os
import re
import shutil
import subprocess

ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
GO, PR = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__cluster_size", "launch__occupancy_limit", "launch__shared_mem_per_block", "sm__warps_active.avg.pct",
        "smsp__issue_active.avg.pct", "sm__pipe_fma_cycles_active.avg.pct", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "gpu__dram_throughput.avg.pct", "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__average_warps_issue_stalled"]


def export(rep, page):
    out = os.path.join(GO, f"{os.path.basename(rep)[:-8]}_{page}.csv")
    with open(out, "w") as f:
        subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], stdout=f, stderr=subprocess.DEVNULL, check=False)
    return list(csv.reader(open(out)))


def key_metrics(rep, dst):
    rows = export(rep, "raw")
    hdr, units = rows[0], rows[1]
    keep = [h for h in hdr if any(t in h for t in WANT) and not h.endswith("peak_sustained") and ".per_second" not in h]
    ki = hdr.index("Kernel Name")
    with open(dst, "w") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit"] + [r[ki] for r in rows[2:]])
        for h in keep:
            i = hdr.index(h)
            w.writerow([h, units[i]] + [r[i] for r in rows[2:]])
    return rows


def dram_bytes(rows, name):
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    mult = {"Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "Gbyte": 1e9}
    names = name if isinstance(name, tuple) else (name,)
    r = [r for r in rows[2:] if any(n in r[ki] for n in names)]
    if not r:
        return None
    tot = 0.0
    for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        i = hdr.index(k)
        tot += float(r[0][i]) * mult[units[i]]
    return tot


traffic = {}
b16 = os.path.join(GO, "prof_r2_batch16.ncu-rep")
big = os.path.join(GO, "prof_r2_chipfull.ncu-rep")
if os.path.exists(b16):
    rows = key_metrics(b16, os.path.join(PR, "r2_ncu_full_key_metrics.csv"))
    # a lone batch-16 call takes the cluster form of the page kernel, the 29-batch launch the one-CTA form
    for key, name in (("als_sparse", ("als_pages_kernel", "als_pages_cluster_kernel")), ("als_sparsify", "als_sparsify_raw_kernel"),
                      ("als_dense", "als_kernel"), ("fuse_tail", "fuse_tail_kernel")):
        traffic[f"{key}_dram_bytes_per_launch"] = dram_bytes(rows, name)
if os.path.exists(big):
    rows = key_metrics(big, os.path.join(PR, "r2_chip_full_key_metrics.csv"))
    for key, name in (("als_sparse", "als_pages_kernel"), ("als_sparsify", "als_sparsify_raw_kernel"), ("als_dense", "als_kernel"), ("fuse_tail", "fuse_tail_kernel")):
        traffic[f"{key}_grouped_dram_bytes_per_launch"] = dram_bytes(rows, name)
    traffic["grouped_batches_per_launch_in_capture"] = 29
    # stall summary of the dominant kernel
    src = export(big, "source")
    secs = [i for i, r in enumerate(src) if r and r[0] == "Kernel Name"]
    st = [i for i in secs if "als_pages_kernel" in src[i][1]][0]
    hdr = src[st + 1]
    body = src[st + 2:min([i for i in secs if i > st] + [len(src)])]
    idx = {h: i for i, h in enumerate(hdr)}
    stall = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    ns = sum(int(r[idx["# Samples"]]) for r in body if r[idx["# Samples"]].isdigit())
    out = [f"kernel: {src[st][1]}", f"warp-state samples: {ns}, SASS instructions: {len(body)}"]
    totals = {h: sum(int(r[idx[h]]) for r in body if r[idx[h]].isdigit()) for h in stall}
    for h, v in sorted(totals.items(), key=lambda kv: -kv[1])[:10]:
        out.append(f"  {h:28s} {v:6d} {100 * v / ns:5.1f}%")
    cnt = collections.Counter(int(r[idx["Instructions Executed"]]) for r in body if r[idx["Instructions Executed"]].isdigit())
    out.append("instructions by execution count (count x static instructions): " + str(sorted(cnt.items(), key=lambda kv: -kv[0] * kv[1])[:5]))
    mx = max(cnt)
    loop = [r for r in body if r[idx["Instructions Executed"]].isdigit() and int(r[idx["Instructions Executed"]]) >= 0.9 * mx]
    out.append(f"iteration loop: {len(loop)} instructions, {sum(int(r[idx['# Samples']]) for r in loop)} samples; hottest instructions:")
    for r in sorted(loop, key=lambda r: -int(r[idx["# Samples"]]))[:25]:
        out.append(f"  {int(r[idx['# Samples']]):4d}  {r[idx['Source']].strip()[:84]:84s} "
                   + str({h[6:]: int(r[idx[h]]) for h in stall if r[idx[h]] not in ("0", "") and int(r[idx[h]]) > 5}))
    open(os.path.join(PR, "r2_als_pages_kernel_stalls.txt"), "w").write("\n".join(out) + "\n")
    print("\n".join(out[:16]))
traffic["source"] = ("profiles/r2_ncu_full_key_metrics.csv (one batch-16 call) and profiles/r2_chip_full_key_metrics.csv (29 batches per launch), "
                     "ncu --set full --clock-control none, raw-matrix inputs; dram__bytes_read.sum + dram__bytes_write.sum per launch")
json.dump(traffic, open(os.path.join(PR, "traffic.json"), "w"), indent=1)

lp = os.path.join(GO, "launches_r2.csv")
if os.path.exists(lp):
    shutil.copy(lp, os.path.join(PR, "r2_launches_bench.csv"))
    rows = [r for r in csv.reader(open(lp)) if r and not r[0].startswith("==")]
    idx = {h: i for i, h in enumerate(rows[0])}
    agg = collections.defaultdict(list)
    for r in rows[1:]:
        try:
            v, u = float(r[idx["Metric Value"]]), r[idx["Metric Unit"]]
        except Exception:
            continue
        agg[re.sub(r"\(.*", "", r[idx["Kernel Name"]])].append(v / 1e3 if u == "ns" else v)
    tot = sum(sum(v) for v in agg.values())
    with open(os.path.join(PR, "r2_launches_summary.csv"), "w") as f:
        f.write("kernel,launches,avg_us,share_pct\n")
        for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            f.write(f"\"{k}\",{len(v)},{sum(v) / len(v):.2f},{100 * sum(v) / tot:.1f}\n")
    print(open(os.path.join(PR, "r2_launches_summary.csv")).read())

# SASS evidence: bulk async copy + mbarrier in the sparsify kernel, named barriers in the pages kernel, cluster barrier in the dense kernel
sass = subprocess.run(["cuobjdump", "-sass", os.path.join(ROOT, "md_rdm_b200", "librdm_b200.so")], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", sass)
ex = []
for name, pats in (("als_sparsify_raw_kernel", r"UBLKCP|SYNCS|MBARRIER"), ("als_pages_kernel", r"BAR\.(ARV|SYNC)"),
                   ("als_pages_cluster_kernel", r"MAPA|SYNCS|STAS|ST\.ASYNC|UCGABAR|MEMBAR"),
                   ("als_kernelENS", r"UCGABAR|CGABAR|BAR\."), ("conv_head_kernel", r"UCGABAR|MAPA|LD\.E|ST\.E"),
                   ("gt_prepare_kernel", r"UCGABAR|MAPA")):
    for f in funcs:
        if name in f.split("\n")[0]:
            lines = [ln.strip() for ln in f.split("\n") if re.search(pats, ln)]
            ex.append(f"== {f.split(chr(10))[0].strip()}  ({len(lines)} matching instructions)")
            ex += ["   " + re.sub(r"\s*/\* 0x[0-9a-f]+ \*/", "", ln) for ln in lines[:12]]
            break
open(os.path.join(PR, "r2_sass_excerpts.txt"), "w").write("\n".join(ex) + "\n")
print("\n".join(ex[:30]))
