#!/bin/bash
# Round capture on a B200 box (run through gpurun from the repo root):  bash profiles/tools/capture.sh r01i
#   1. the plain bench line (never taken under a profiler)
#   2. the ncu launch list of the same command (per-launch times are cold-cache and serialised: compare shares)
#   3. one `ncu --set full` capture of the three anneal launches of ONE timed step (size classes of equal occupancy are merged into one launch) (the three warm-up steps are skipped)
#   4. device-timed throughput of the other BASELINE configurations
set -x
tag=${1:-rXX}
out=gpurun_out
python bench.py --steps 20 --warmup 3 > $out/${tag}_bench.json 2> $out/${tag}_bench.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu > $out/${tag}_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sats_anneal_kernel --launch-skip 9 --launch-count 3 \
    -f -o $out/prof_${tag} python bench.py --steps 2 --warmup 3 --no-cpu > $out/${tag}_ncu_full.log 2>&1
python profiles/tools/perf_configs.py > $out/${tag}_configs.jsonl 2> $out/${tag}_configs.err
tail -1 $out/${tag}_bench.json | cut -c1-400
