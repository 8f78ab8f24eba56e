#!/usr/bin/env python
"""One run of the unmodified reference `-c` CPU path over the WHOLE bench database (BASELINE configs[4]: D2PHLB1 vs the
100 000-structure synthetic db, 128 restarts), P processes on P interleaved shards of the size-sorted db (the reference is
single threaded with a process-global drand48 stream), P = all host cores.  Pins the rate that bench.py's reference arm
extrapolates from a 600-structure-per-core sample.  Run on the GPU box's host:

  python profiles/tools/reference_full.py > gpurun_out/r02_reference_full100k.json      (then copy into profiles/)"""
import json
import os
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
import cuda_satabsearch_b200 as S  # noqa: E402   (only to write the 148 MB of ASCII input quickly)

db = bench.synthetic_db(S)
qs = bench.query_db(S)
cores = bench.host_cores()
with tempfile.TemporaryDirectory() as td:
    qs.write_ascii(os.path.join(td, "q.ascii"))
    qtext = Path(td, "q.ascii").read_text()
    n = len(db)
    for p in range(cores):
        db.select(np.arange(p, n, cores, dtype=np.int32)).write_ascii(os.path.join(td, "db%d.ascii" % p))
        Path(td, "in%d" % p).write_text("db%d.ascii\nT T F\n%s" % (p, qtext))
    t0 = time.perf_counter()
    v, det = bench._reference_processes(td, cores, n / cores)
    wall = time.perf_counter() - t0
print(json.dumps({"what": "unmodified reference -c path, whole 100k synthetic db, D2PHLB1, 128 restarts", "cores": cores,
                  "structures": n, "structures_per_s": v, "slowest_process_search_s": det["slowest_search_s"],
                  "wall_s_incl_parse": wall, "core_seconds_search": det["slowest_search_s"] * cores}))
