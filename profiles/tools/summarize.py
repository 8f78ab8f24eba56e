#!/usr/bin/env python
"""Summarise ncu outputs brought back in gpurun_out/ into small text files under profiles/.

  python profiles/tools/summarize.py launches gpurun_out/launches.csv            > profiles/rNN_launches.txt
  python profiles/tools/summarize.py raw      gpurun_out/prof.ncu-rep            > profiles/rNN_kernels.txt
  python profiles/tools/summarize.py source   gpurun_out/prof.ncu-rep <kernel#>  > profiles/rNN_source_kK.txt
"""
import collections
import csv
import io
import re
import subprocess
import sys

RAW = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
       "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
       "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
       "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
       "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
       "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
       "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
       "derived__memory_l1_wavefronts_shared_excessive", "l1tex__lsuin_requests.avg.pct_of_peak_sustained_elapsed",
       "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__cycles_elapsed.max"]


def launches(path):
    rows = list(csv.reader(l for l in open(path) if not l.startswith("==")))
    h = rows[0]
    ki, vi, gi, bi = h.index("Kernel Name"), h.index("Metric Value"), h.index("Grid Size"), h.index("Block Size")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        if len(r) > vi:
            agg.setdefault((re.sub(r"\(.*", "", r[ki])[:64], r[gi], r[bi]), []).append(float(r[vi].replace(",", "")))
    tot = sum(sum(v) for v in agg.values())
    print("# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)")
    for k, v in agg.items():
        print("%-66s grid=%-14s block=%-13s n=%3d mean=%9.1f us share=%5.1f%%" % (k[0], k[1], k[2], len(v), sum(v) / len(v) / 1e3, 100 * sum(v) / tot))


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    h, u = rows[0], rows[1]
    idx = {n: i for i, n in enumerate(h)}
    print("# ncu --set full --clock-control none; one column per captured launch")
    for w in ["Kernel Name", "Grid Size", "Block Size"] + RAW:
        if w in idx:
            print("%-78s %-16s %s" % (w, u[idx[w]], " | ".join(r[idx[w]][:26] for r in rows[2:])))


def source(rep, kid, top=60):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-id", ":::" + kid],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    h = [i for i, r in enumerate(rows) if r and r[0] == "Line No"][0]
    fn = [r[1] for r in rows[:h] if r and r[0] == "Function Name"]
    src = {}
    agg = collections.defaultdict(lambda: [0, 0, 0])
    tot = [0, 0, 0]
    for r in rows[h + 1:]:
        if len(r) < 10 or not r[0].isdigit():
            continue
        try:
            v = (int(r[7]), int(r[8]), int(r[6]))
        except ValueError:
            continue
        src[int(r[0])] = r[1]
        for k in range(3):
            agg[int(r[0])][k] += v[k]
            tot[k] += v[k]
    print("# %s" % (fn[0] if fn else ""))
    print("# warp instructions executed: %d ; thread instructions: %d ; avg active threads/instruction: %.2f" % (tot[0], tot[1], tot[1] / max(tot[0], 1)))
    print("# line  share-of-warp-inst  active-threads/inst  share-of-stall-samples  source")
    for ln, (ie, te, sm) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print("%4d  %5.1f%%  %4.1f  %5.1f%%  %s" % (ln, 100 * ie / tot[0], te / max(ie, 1), 100 * sm / max(tot[2], 1), src.get(ln, "").strip()[:120]))


if __name__ == "__main__":
    {"launches": launches, "raw": raw, "source": source}[sys.argv[1]](*sys.argv[2:])
