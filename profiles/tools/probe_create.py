#!/usr/bin/env python
"""Where does multi-GPU start-up time go?  Times CUDA context creation and searcher creation per device, sequentially and
from threads (what `cudaSaTabsearch -g N` does)."""
import ctypes as C
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
import cuda_satabsearch_b200 as S  # noqa: E402

rt = C.CDLL("libcudart.so.12")
n = S.device_count()
base = S.Database.read_packed(ROOT / "tests/golden/small586.satsdb")
db = base.bootstrap(100000, 20240502, True)
mode = sys.argv[1] if len(sys.argv) > 1 else "seq"


def ctx(dev):
    t0 = time.perf_counter()
    rt.cudaSetDevice(dev)
    rt.cudaFree(None)
    return (time.perf_counter() - t0) * 1e3


def create(dev, rank, count, out):
    t0 = time.perf_counter()
    sr = S.Searcher(db, dev, rank, count)
    out[rank] = ((time.perf_counter() - t0) * 1e3, sr)


if mode == "seq":
    for d in range(n):
        print("context dev %d: %.1f ms" % (d, ctx(d)), flush=True)
    out = {}
    for d in range(n):
        create(d, d, n, out)
        print("searcher shard %d/%d on dev %d: %.1f ms" % (d, n, d, out[d][0]), flush=True)
else:
    out = {}
    t0 = time.perf_counter()
    th = [threading.Thread(target=create, args=(d, d, n, out)) for d in range(n)]
    [t.start() for t in th]
    [t.join() for t in th]
    print("threads: total %.1f ms; per shard %s" % ((time.perf_counter() - t0) * 1e3, [round(out[d][0], 1) for d in range(n)]), flush=True)
p = S.default_params(restarts=128)
qs = S.Database.read_packed(ROOT / "tests/golden/queries.satsdb")
q = qs.select([qs.find("D2PHLB1")])
for rep in range(3):
    for d in range(n):
        t0 = time.perf_counter()
        out[d][1].upload(q)
        out[d][1].launch(p, 0)
        out[d][1].sync()
        print("rep %d dev %d upload+launch+sync %.1f ms" % (rep, d, (time.perf_counter() - t0) * 1e3), flush=True)
