#!/usr/bin/env python
"""BASELINE configs[2] on several GPUs: 200 queries (drawn from the synthetic 14 297-structure db, seed 200) against that db,
R = 128, one searcher per GPU in ONE process (what `cudaSaTabsearch -q db -g N` does): every GPU holds a cost-weighted shard of
the db and searches all queries against it; launches are asynchronous, so the GPUs run side by side.

  python profiles/tools/multi_query_scaling.py [N ...]        one JSON line per N (default: 1 and every GPU present)"""
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
import cuda_satabsearch_b200 as S  # noqa: E402

have = S.device_count()
ns = [int(a) for a in sys.argv[1:]] or sorted({1, have})
base = S.Database.read_packed(ROOT / "tests/golden/small586.satsdb")
db = base.bootstrap(14297, 20240501, True)
rng = np.random.default_rng(200)
queries = db.select(rng.choice(len(db), 200, replace=False).astype(np.int32))
p = S.default_params(lorder=1, lsoln=0, restarts=128, seed=4242)
ref = None
for n in ns:
    srs = [S.Searcher(db, g % have, g, n) for g in range(n)]
    scores = np.full((len(queries), len(db)), np.iinfo(np.int32).min, np.int32)

    def step():
        for sr in srs:
            sr.upload(queries)
        for sr in srs:
            sr.launch(p, 0)
        for sr in srs:
            sr.collect_begin()
        for sr in srs:
            sr.collect(scores=scores)

    for _ in range(2):
        step()
    walls = []
    for _ in range(3):
        t0 = time.perf_counter()
        step()
        walls.append((time.perf_counter() - t0) * 1e3)
    dev = [max(sr.launch(p, 0, timed=True) for _ in range(2)) for sr in srs]       # kernels only, one GPU at a time
    if ref is None:
        ref = scores.copy()
    pairs = len(queries) * len(db)
    print(json.dumps({"config": "200 queries (-q mode) TTF vs 14297, R=128", "gpus": n, "gpus_present": have,
                      "wall_ms_upload_launch_collect": float(np.median(walls)), "device_ms_per_gpu": dev,
                      "pairs_per_s_wall": pairs / (float(np.median(walls)) / 1e3),
                      "d2h_bytes": int(4 * pairs), "identical_to_first": bool(np.array_equal(scores, ref))}), flush=True)
    for sr in srs:
        sr.close()
