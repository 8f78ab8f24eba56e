#!/usr/bin/env python
"""A small tour of every device code path (production and validation kernels, 1/2/4-word masks, LSOLN, multi-query batches,
top-k, significance cut, streaming hits) for compute-sanitizer:

  compute-sanitizer --tool memcheck  python profiles/tools/sanitize.py
  compute-sanitizer --tool racecheck python profiles/tools/sanitize.py
  compute-sanitizer --tool synccheck python profiles/tools/sanitize.py"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import cuda_satabsearch_b200 as S  # noqa: E402
from _refio import GOLDEN, read_packed  # noqa: E402

ents = read_packed(GOLDEN / "small586.satsdb")
big = sorted(ents, key=lambda s: -s.n)[:6]
pick = ents[:40] + big
qs = {s.name: s for s in read_packed(GOLDEN / "queries.satsdb")}
db = S.Database.from_structures([s.name for s in pick], [s.tab for s in pick], [s.dmat for s in pick])
queries = [qs[n] for n in ("D1UBIA_", "D2PHLB1", "SHEETBC", "d1twfa_")]
qdb = S.Database.from_structures([q.name for q in queries], [q.tab for q in queries], [q.dmat for q in queries])
sr = S.Searcher(db, 0)
total = 0
for lorder in (1, 0):
    for lsoln in (0, 1):
        sc, mp = sr.search(qdb, S.default_params(lorder=lorder, lsoln=lsoln, restarts=40, seed=3))
        total += int(sc.sum())
sc, mp = sr.search(qdb, S.default_params(lorder=1, lsoln=1, restarts=128, rng_mode=S.RNG_XORWOW_GRID), qcount=2)
total += int(sc.sum())
sr.upload(qdb)
sr.bind_cut(0.5)
sr.launch(S.default_params(restarts=33))
cnt, idx, hs, nbytes = sr.streamed_hits(50)
total += int(cnt.sum())
cnt2, _, _ = sr.hits(0.5, 50)
assert (cnt == cnt2).all()
idx, tsc = sr.topk(5)
total += int(tsc.sum())
sr.bind_cut(None)
sr.close()
shard = S.Searcher(db, 0, 1, 3)
sc, _ = shard.search(qdb, S.default_params(restarts=32))
shard.close()
print("sanitize tour done, checksum", total)
