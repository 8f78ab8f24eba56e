#!/bin/bash
# ncu --set full capture of ONE of perf_configs.py's configurations (all anneal launches of its last timed search):
#   bash profiles/tools/capture_config.sh r02h_sheetbc SHEETBC <launches-per-search>
# perf_configs.py --reps 1 runs 2 warm-up searches + 1 timed one: the last <launches-per-search> anneal launches are kept.
tag=$1; only=$2; per=${3:-4}
out=gpurun_out
python profiles/tools/perf_configs.py --only "$only" --reps 1 > $out/${tag}_time.json 2> $out/${tag}_time.err || exit 1
ncu --set full --clock-control none --import-source on -k regex:sats_anneal_kernel --launch-count 40 \
    -f -o $out/prof_${tag} python profiles/tools/perf_configs.py --only "$only" --reps 1 > $out/${tag}_ncu.log 2>&1
cat $out/${tag}_time.json
