#!/usr/bin/env python
"""Host-timed cost of one complete search call (upload + launch + collect) on a 1/8 shard and on the whole bench db, per
libsats build: what the launch-plan cache and other host-side work show up in.  python profiles/tools/e2e_time.py a=lib.so b=lib.so"""
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent))
from ab_time import Lib, ROOT  # noqa: E402

for a in sys.argv[1:]:
    label, path = a.split("=", 1)
    lib = Lib(path)
    lib.L.sats_search.argtypes = [__import__("ctypes").c_void_p] * 2 + [__import__("ctypes").c_int] * 2 + [__import__("ctypes").c_void_p, __import__("ctypes").c_uint32] + [__import__("ctypes").c_void_p] * 2
    base = lib.read_packed(ROOT / "tests/golden/small586.satsdb")
    qs = lib.read_packed(ROOT / "tests/golden/queries.satsdb")
    db = lib.bootstrap(base, 100000, 20240502)
    q = lib.query(qs, "D2PHLB1")
    for shards in (1, 8):
        sr = lib.searcher(db, 0, shards)
        p = lib.params(restarts=128)
        sc = np.zeros((1, 100000), np.int32)
        import ctypes as C
        for _ in range(3):
            lib.ck(lib.L.sats_search(sr, q, 0, 1, C.byref(p), 0, sc.ctypes.data, None))
        ts = []
        for _ in range(30):
            t0 = time.perf_counter()
            lib.ck(lib.L.sats_search(sr, q, 0, 1, C.byref(p), 0, sc.ctypes.data, None))
            ts.append((time.perf_counter() - t0) * 1e3)
        dev = float(np.median([lib.launch(sr, p) for _ in range(7)]))
        print("%-8s shards=%d  sats_search wall median %.4f ms (min %.4f)   launch device-timed %.4f ms" % (label, shards, float(np.median(ts)), min(ts), dev), flush=True)
        lib.L.sats_searcher_free(sr)
