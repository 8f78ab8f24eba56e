#!/usr/bin/env python
"""Fold one `ncu --set full` capture of a bench step (the anneal launches of ONE step) into profiles/latest.json,
the numbers bench.py quotes in its roofline object.

  python profiles/tools/make_latest.py gpurun_out/prof_r1i.ncu-rep "profiles/r01i_* (...)" [move_evals_per_step]
"""
import csv
import io
import json
import subprocess
import sys
from pathlib import Path

rep, capture = sys.argv[1], sys.argv[2]
moves = float(sys.argv[3]) if len(sys.argv) > 3 else 100000 * 128 * 100.0
keep = int(sys.argv[4]) if len(sys.argv) > 4 else None          # launches of ONE step (the first `keep` captured ones)
if rep.endswith(".csv"):                                         # an `ncu --page raw --csv` export instead of the report itself
    out = open(rep).read()
else:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
if keep:
    rows = rows[:2 + keep]
h = {n: i for i, n in enumerate(rows[0])}
units = rows[1]


def col(name):
    return [float(r[h[name]].replace(",", "")) for r in rows[2:]]


def scaled(name, want):
    """ncu picks a unit per column (e.g. Mbyte); bring it to bytes / ms."""
    u = units[h[name]].lower()
    f = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[u]
    return [v * f for v in col(name)]


inst = col("smsp__inst_executed.sum")
ratio = col("smsp__thread_inst_executed_per_inst_executed.ratio")
t = scaled("gpu__time_duration.sum", "ms")
issue = col("smsp__issue_active.avg.pct_of_peak_sustained_active")
smem = col("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed")
dram = [a + b for a, b in zip(scaled("dram__bytes_read.sum", "byte"), scaled("dram__bytes_write.sum", "byte"))]
pipes = {p: col("sm__inst_executed_pipe_%s.avg.pct_of_peak_sustained_active" % p) for p in ("alu", "fma", "xu", "lsu")
         if "sm__inst_executed_pipe_%s.avg.pct_of_peak_sustained_active" % p in h}
T = sum(t)
res = {
    "capture": capture,
    "launches": len(inst),
    "warp_inst_per_step": sum(inst),
    "warp_inst_per_move": sum(inst) / moves,
    "avg_active_threads_per_inst": sum(a * b for a, b in zip(inst, ratio)) / sum(inst),
    "issue_active_pct_time_weighted": sum(a * b for a, b in zip(issue, t)) / T,
    "smem_wavefront_pct_of_peak_time_weighted": sum(a * b for a, b in zip(smem, t)) / T,
    "pipe_pct_time_weighted": {p: sum(a * b for a, b in zip(v, t)) / T for p, v in pipes.items()},
    "dram_bytes_per_step": sum(dram),
    "sum_kernel_ms": T,
}
Path(__file__).resolve().parents[1].joinpath("latest.json").write_text(json.dumps(res, indent=1) + "\n")
print(json.dumps(res, indent=1))
