#!/bin/bash
# Everything a round's evidence needs, summarised ON THE BOX (ncu reports are too large to bring back in numbers):
#   bash profiles/tools/capture_all.sh r02g      -> gpurun_out/r02g_*.{json,txt,jsonl}
tag=${1:-rXX}
out=gpurun_out
bash profiles/tools/capture.sh $tag > $out/${tag}_capture.log 2>&1
python profiles/tools/summarize.py launches $out/${tag}_launches.csv > $out/${tag}_launches.txt
python profiles/tools/summarize.py raw $out/prof_${tag}.ncu-rep > $out/${tag}_kernels.txt
python profiles/tools/summarize.py source $out/prof_${tag}.ncu-rep 3 > $out/${tag}_source_k3.txt
python profiles/tools/make_latest.py $out/prof_${tag}.ncu-rep "profiles/${tag}_* (ncu --set full, bench.py --steps 2 --warmup 3 --no-cpu, one B200, 3 launches = one step)" 1.28e9 3 > $out/${tag}_latest.json
ncu -i $out/prof_${tag}.ncu-rep --page raw --csv > $out/${tag}_raw.csv 2>/dev/null
rm -f $out/prof_${tag}.ncu-rep
for cfg in "sheetbc SHEETBC 165" "n101 d1twfa_ 0" "ubia D1UBIA_ 69"; do
  set -- $cfg
  bash profiles/tools/capture_config.sh ${tag}_$1 $2 > $out/${tag}_$1.log 2>&1
  python profiles/tools/summarize.py raw $out/prof_${tag}_$1.ncu-rep > $out/${tag}_$1_kernels.txt
  python - $out/prof_${tag}_$1.ncu-rep $out/${tag}_$1_time.json <<'PY' > $out/${tag}_$1_summary.json
import csv, io, json, subprocess, sys
rep, timejson = sys.argv[1], sys.argv[2]
rows = list(csv.reader(io.StringIO(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout)))
h = {n: i for i, n in enumerate(rows[0])}
def col(name): return [float(r[h[name]].replace(",", "")) for r in rows[2:]]
inst, ratio, issue = col("smsp__inst_executed.sum"), col("smsp__thread_inst_executed_per_inst_executed.ratio"), col("smsp__issue_active.avg.pct_of_peak_sustained_active")
smem = col("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed")
unit = rows[1][h["gpu__time_duration.sum"]]
t = [v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[unit] for v in col("gpu__time_duration.sum")]
n = len(inst) // 3                      # perf_configs.py --reps 1 = 2 warm-up searches + 1 timed: keep the last third
sl = slice(len(inst) - n, len(inst))
tm = json.loads(open(timejson).read().strip().split("\n")[-1])
moves = tm["move_evals_per_s"] * tm["ms"] / 1e3
T = sum(t[sl])
wi = sum(inst[sl])
print(json.dumps({"config": tm["config"], "ms_device_timed": tm["ms"], "move_evals_per_s": tm["move_evals_per_s"], "launches_per_search": n,
                  "warp_inst_per_search": wi, "warp_inst_per_move": wi / moves,
                  "avg_active_threads_per_inst": sum(a * b for a, b in zip(inst[sl], ratio[sl])) / wi,
                  "issue_active_pct_time_weighted": sum(a * b for a, b in zip(issue[sl], t[sl])) / T,
                  "smem_wavefront_pct_of_peak_time_weighted": sum(a * b for a, b in zip(smem[sl], t[sl])) / T,
                  "issue_roofline_frac": tm["move_evals_per_s"] * (wi / moves) / tm["issue_peak_warp_inst_per_s"],
                  "smem_roofline": tm.get("smem_roofline"), "sum_kernel_ms_under_ncu": T}))
PY
  rm -f $out/prof_${tag}_$1.ncu-rep
done
ls -la $out
