#!/usr/bin/env python
"""Same-box A/B timing of libsats builds (the tool behind the kernel experiments in profiles/README.md).

  python profiles/tools/ab_time.py label=path/to/libsats.so [label2=other.so ...] [--reps N] [--only bench,lsoln,...]

Binds the handful of C-ABI entry points it needs by hand (so that older builds, which lack newer symbols, load too),
runs every workload on every build in turn -- interleaved, so that clock drift hits all builds alike -- and prints the
device-timed median per (build, workload) plus a checksum of the scores (builds with the same stream layout must agree;
parity proper is tests/ -m gpu)."""
import ctypes as C
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]


class Params(C.Structure):
    _fields_ = [("lorder", C.c_int), ("lsoln", C.c_int), ("restarts", C.c_int), ("rng_mode", C.c_int),
                ("accept_mode", C.c_int), ("pool", C.c_int), ("pool_threshold", C.c_int),
                ("grid_rank", C.c_int), ("grid_count", C.c_int), ("reserved", C.c_int), ("seed", C.c_uint64)]


class Lib:
    def __init__(self, path):
        L = self.L = C.CDLL(str(path))
        vp, ci = C.c_void_p, C.c_int
        P = C.POINTER
        L.sats_last_error.restype = C.c_char_p
        L.sats_db_read_packed.argtypes = [C.c_char_p, P(vp)]
        L.sats_db_bootstrap.argtypes = [vp, ci, C.c_uint64, ci, P(vp)]
        L.sats_db_select.argtypes = [vp, vp, ci, P(vp)]
        L.sats_db_find.argtypes = [vp, C.c_char_p]
        L.sats_db_count.argtypes = [vp]
        L.sats_params_default.argtypes = [P(Params)]
        L.sats_searcher_create.argtypes = [vp, ci, ci, ci, P(vp)]
        L.sats_searcher_free.argtypes = [vp]
        L.sats_search_upload.argtypes = [vp, vp, ci, ci]
        L.sats_search_launch.argtypes = [vp, P(Params), C.c_uint32, P(C.c_float)]
        L.sats_search_collect.argtypes = [vp, vp, vp]

    def ck(self, rc):
        if rc < 0:
            raise RuntimeError(self.L.sats_last_error().decode())
        return rc

    def read_packed(self, path):
        h = C.c_void_p()
        self.ck(self.L.sats_db_read_packed(str(path).encode(), C.byref(h)))
        return h

    def bootstrap(self, db, n, seed):
        h = C.c_void_p()
        self.ck(self.L.sats_db_bootstrap(db, n, seed, 1, C.byref(h)))
        return h

    def query(self, qs, name):
        idx = np.array([self.ck(self.L.sats_db_find(qs, name.encode()))], np.int32)
        h = C.c_void_p()
        self.ck(self.L.sats_db_select(qs, idx.ctypes.data, 1, C.byref(h)))
        return h

    def params(self, **kw):
        p = Params()
        self.L.sats_params_default(C.byref(p))
        for k, v in kw.items():
            setattr(p, k, int(v))
        return p

    def searcher(self, db, rank=0, count=1):
        h = C.c_void_p()
        self.ck(self.L.sats_searcher_create(db, 0, rank, count, C.byref(h)))
        return h

    def launch(self, sr, p):
        ms = C.c_float(0)
        self.ck(self.L.sats_search_launch(sr, C.byref(p), 0, C.byref(ms)))
        return ms.value


WORKLOADS = {
    # key: (description, db, query, shards, params)
    "bench": ("D2PHLB1 TTF vs 100k, R=128 (bench workload)", "db100", "D2PHLB1", 1, dict(restarts=128)),
    "lsoln": ("the same, LSOLN=T", "db100", "D2PHLB1", 1, dict(restarts=128, lsoln=1)),
    "shard8": ("the same on a 1/8 shard", "db100", "D2PHLB1", 8, dict(restarts=128)),
    "sheet": ("SHEETBC TFT vs 14297, R=1024", "db15", "SHEETBC", 1, dict(restarts=1024, lorder=0, lsoln=1)),
    "r1024": ("D2PHLB1 TTF vs 14297, R=1024", "db15", "D2PHLB1", 1, dict(restarts=1024)),
    "n101": ("d1twfa_ (n1=101) TTF vs 14297, R=128", "db15", "d1twfa_", 1, dict(restarts=128)),
    "ubia": ("D1UBIA_ (n1=8) TTF vs 100k, R=128", "db100", "D1UBIA_", 1, dict(restarts=128)),
}


def main():
    args = [a for a in sys.argv[1:] if "=" in a and not a.startswith("--")]
    reps = 5
    only = list(WORKLOADS)
    for i, a in enumerate(sys.argv):
        if a == "--reps":
            reps = int(sys.argv[i + 1])
        if a == "--only":
            only = sys.argv[i + 1].split(",")
    builds = []
    for a in args:
        label, path = a.split("=", 1)
        lib = Lib(path)
        base = lib.read_packed(ROOT / "tests/golden/small586.satsdb")
        qs = lib.read_packed(ROOT / "tests/golden/queries.satsdb")
        dbs = {"db100": lib.bootstrap(base, 100000, 20240502), "db15": lib.bootstrap(base, 14297, 20240501)}
        builds.append((label, lib, qs, dbs))
    for key in only:
        what, dbname, qname, shards, kw = WORKLOADS[key]
        state = []
        for label, lib, qs, dbs in builds:
            sr = lib.searcher(dbs[dbname], 0, shards)
            p = lib.params(**kw)
            lib.ck(lib.L.sats_search_upload(sr, lib.query(qs, qname), 0, 1))
            for _ in range(2):
                lib.launch(sr, p)
            state.append((label, lib, sr, p, dbs[dbname], []))
        for _ in range(reps):                      # interleave the builds
            for label, lib, sr, p, db, times in state:
                times.append(lib.launch(sr, p))
        for label, lib, sr, p, db, times in state:
            n = lib.L.sats_db_count(db)
            sc = np.zeros((1, n), np.int32)
            mp = np.zeros((1, n, 111), np.int32) if p.lsoln else None
            lib.ck(lib.L.sats_search_collect(sr, sc.ctypes.data, mp.ctypes.data if p.lsoln else None))
            print("%-10s %-46s %8.3f ms  (min %8.3f)  checksum %d" % (label, what, float(np.median(times)), min(times),
                                                                     int(sc.astype(np.int64).sum())), flush=True)
            lib.L.sats_searcher_free(sr)


if __name__ == "__main__":
    main()
