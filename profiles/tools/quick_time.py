#!/usr/bin/env python
"""Same-box A/B timing used for the kernel experiments in profiles/README.md: device-timed median of a few searches per
configuration, plus a checksum of the scores (a cheap "results did not change" probe; parity proper is tests/ -m gpu).

  python profiles/tools/quick_time.py [label]           # on a B200 box, from the repo root
Swap cuda_satabsearch_b200/libsats.so between runs to compare builds (make ... LIBOUT=/some/other.so EXTRA=-DFLAG)."""
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
import cuda_satabsearch_b200 as S  # noqa: E402

label = sys.argv[1] if len(sys.argv) > 1 else "build"
base = S.Database.read_packed(ROOT / "tests/golden/small586.satsdb")
qs = S.Database.read_packed(ROOT / "tests/golden/queries.satsdb")
db100 = base.bootstrap(100000, 20240502, True)
db15 = base.bootstrap(14297, 20240501, True)


def q(name):
    return qs.select([qs.find(name)])


def run(what, db, queries, reps=5, shards=1, **kw):
    sr = S.Searcher(db, 0, 0, shards)
    p = S.default_params(**kw)
    sr.upload(queries)
    for _ in range(2):
        sr.launch(p, 0, timed=True)
    ms = float(np.median([sr.launch(p, 0, timed=True) for _ in range(reps)]))
    sc, _ = sr.collect()
    t0 = time.perf_counter()
    sr.search(queries, p)
    e2e = (time.perf_counter() - t0) * 1e3
    print("%-12s %-44s %8.3f ms  (one sats_search: %8.3f ms)  checksum %d" % (label, what, ms, e2e, int(sc.astype(np.int64).sum())), flush=True)
    sr.close()


run("bench workload: D2PHLB1 TTF vs 100k, R=128", db100, q("D2PHLB1"), restarts=128)
run("the same, LSOLN=T", db100, q("D2PHLB1"), restarts=128, lsoln=1)
run("the same on a 1/8 shard", db100, q("D2PHLB1"), restarts=128, shards=8)
run("SHEETBC TFT vs 14297, R=1024", db15, q("SHEETBC"), reps=3, restarts=1024, lorder=0, lsoln=1)
run("D2PHLB1 TTF vs 14297, R=1024", db15, q("D2PHLB1"), reps=3, restarts=1024)
run("d1twfa_ (n1=101) TTF vs 14297, R=128", db15, q("d1twfa_"), reps=3, restarts=128)
