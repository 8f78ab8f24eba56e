#!/usr/bin/env python
"""Device-timed throughput of the BASELINE configurations other than the bench workload (run on a B200).

  python profiles/tools/perf_configs.py                       every configuration, one JSON line each
  python profiles/tools/perf_configs.py --only SHEETBC --reps 2   one of them (what capture_config.sh runs under ncu)

Every line carries the shared-memory roofline view where SURVEY 8(d) gives the reference-width bytes per move-eval
(69 / 148 / 165 B for D1UBIA_ / D2PHLB1 / SHEETBC): achieved = move-evals/s x bytes over 148 SMs x 128 B/clk x 1965 MHz."""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
import cuda_satabsearch_b200 as S  # noqa: E402

base = S.Database.read_packed(ROOT / "tests/golden/small586.satsdb")
qs = S.Database.read_packed(ROOT / "tests/golden/queries.satsdb")
db15 = base.bootstrap(14297, 20240501, True)
db100 = base.bootstrap(100000, 20240502, True)


def q(name):
    return qs.select([qs.find(name)])


ONLY = sys.argv[sys.argv.index("--only") + 1] if "--only" in sys.argv else None
REPS = int(sys.argv[sys.argv.index("--reps") + 1]) if "--reps" in sys.argv else None
SMEM_PEAK = 148 * 128 * 1965e6          # bytes/s at the maximum SM clock
ISSUE_PEAK = 148 * 4 * 1965e6           # warp instructions/s


def run(label, db, queries, reps=5, algo_bytes=None, **kw):
    if ONLY and ONLY not in label:
        return None
    reps = REPS or reps
    sr = S.Searcher(db, 0)
    p = S.default_params(**kw)
    sr.upload(queries)
    for _ in range(2):
        sr.launch(p, 0, timed=True)
    ms = float(np.median([sr.launch(p, 0, timed=True) for _ in range(reps)]))
    pairs = len(queries) * len(db)
    moves = pairs * kw.get("restarts", 128) * 100 / ms * 1e3
    out = {"config": label, "ms": ms, "structures_per_s": pairs / ms * 1e3, "move_evals_per_s": moves,
           "smem_roofline": None if algo_bytes is None else {"algorithmic_bytes_per_move": algo_bytes, "achieved_gbs": moves * algo_bytes / 1e9,
                                                              "peak_gbs": SMEM_PEAK / 1e9, "frac": moves * algo_bytes / SMEM_PEAK},
           "issue_peak_warp_inst_per_s": ISSUE_PEAK}
    print(json.dumps(out), flush=True)
    sr.close()
    return out


rng = np.random.default_rng(200)
q200 = db15.select(rng.choice(len(db15), 200, replace=False).astype(np.int32))
run("D2PHLB1 n1=19 TTF vs 100k, R=128 (bench workload)", db100, q("D2PHLB1"), algo_bytes=148.0, restarts=128)
run("D1UBIA_ n1=8 TTF vs 100k, R=128", db100, q("D1UBIA_"), algo_bytes=69.0, restarts=128)
run("D1UBIA_ n1=8 TTT vs 14297, R=128, validation (XORWOW grid) mode", db15, q("D1UBIA_"), reps=3, restarts=128, lsoln=1, rng_mode=S.RNG_XORWOW_GRID)
run("200 queries (-q mode) TTF vs 14297, R=128", db15, q200, reps=3, restarts=128)
run("SHEETBC n1=9 TFT vs 14297, R=1024", db15, q("SHEETBC"), reps=3, algo_bytes=165.0, restarts=1024, lorder=0, lsoln=1)
run("d1twfa_ n1=101 TTF vs 14297, R=128", db15, q("d1twfa_"), reps=3, restarts=128)
run("D2PHLB1 n1=19 TTT vs 100k, R=128 (LSOLN)", db100, q("D2PHLB1"), restarts=128, lsoln=1)
