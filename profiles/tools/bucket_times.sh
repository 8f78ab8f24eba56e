#!/bin/bash
# Per-launch kernel times of one bench step under ncu (serialised, cold cache: compare like with like) for a given libsats build:
#   bash profiles/tools/bucket_times.sh <label> <path/to/libsats.so>
label=$1; lib=$2
cp cuda_satabsearch_b200/libsats.so /tmp/libsats_keep.so
cp "$lib" cuda_satabsearch_b200/libsats.so
ncu --metrics gpu__time_duration.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,launch__shared_mem_per_block_dynamic,launch__grid_size,launch__block_size \
    --clock-control none -k regex:sats_anneal_kernel --launch-skip ${3:-12} --launch-count ${4:-4} --csv --log-file gpurun_out/${label}_buckets.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/${label}_buckets.log 2>&1
cp /tmp/libsats_keep.so cuda_satabsearch_b200/libsats.so
python - "$label" <<'PY'
import csv, sys, collections
rows = list(csv.reader(l for l in open("gpurun_out/%s_buckets.csv" % sys.argv[1]) if not l.startswith("==")))
h = rows[0]
ki, ni, vi, idi = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("ID")
per = collections.OrderedDict()
for r in rows[1:]:
    per.setdefault((r[idi], r[ki][:44]), {})[r[ni]] = r[vi]
for (i, k), m in per.items():
    print(sys.argv[1], i, k, " ".join("%s=%s" % (a.split(".")[0].replace("launch__", "").replace("sm__", "").replace("smsp__", ""), b) for a, b in m.items()))
PY
