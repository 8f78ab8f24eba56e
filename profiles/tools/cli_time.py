#!/usr/bin/env python
"""End-to-end timing of the drop-in binary itself (`cuda_satabsearch_b200/bin/cudaSaTabsearch -g N`), the product's own
multi-GPU path: one process, one searcher per GPU, host-side scatter of the shards' scores.  Run on a B200 box:

  python profiles/tools/cli_time.py [--gpus 1,2,4,8] > gpurun_out/rNN_cli.jsonl

Per run one JSON line: wall clock of the whole process, and the phases the binary reports on stderr -- database load
(packed SATSDB1 cache or ASCII text), searcher creation (blob build + upload, all GPUs side by side), search (upload of
the queries, kernels, collection of every shard: "GPU execution time").  Workloads: BASELINE configs[4] (D2PHLB1 vs the
100k synthetic db) and configs[2] (200 queries drawn from the 14 297-structure synthetic db, -q mode)."""
import json
import os
import re
import subprocess
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
import cuda_satabsearch_b200 as S  # noqa: E402


def run_cli(args, stdin_path, cwd):
    t0 = time.perf_counter()
    with open(stdin_path, "rb") as fh:
        p = subprocess.run([str(S.CLI_PATH)] + [str(a) for a in args], stdin=fh, cwd=cwd, capture_output=True, timeout=1800)
    wall = time.perf_counter() - t0
    err = p.stderr.decode()
    if p.returncode != 0:
        raise RuntimeError(err[-1000:])
    load = re.search(r"Loaded \d+ db entries .* in ([0-9.]+) ms", err)
    copy = re.search(r"Copied \d+ entries to \d+ GPU\(s\) in ([0-9.]+) ms", err)
    search = [float(x) for x in re.findall(r"GPU execution time ([0-9.]+) ms", err)]
    return {"wall_s": wall, "load_ms": float(load.group(1)), "create_ms": float(copy.group(1)), "search_ms": sum(search),
            "search_calls": len(search), "stdout_bytes": len(p.stdout), "stdout_md5": __import__("hashlib").md5(p.stdout).hexdigest()}


def main():
    gpus = [1, 2, 4, 8]
    if "--gpus" in sys.argv:
        gpus = [int(x) for x in sys.argv[sys.argv.index("--gpus") + 1].split(",")]
    have = S.device_count()
    base = S.Database.read_packed(ROOT / "tests/golden/small586.satsdb")
    qs = S.Database.read_packed(ROOT / "tests/golden/queries.satsdb")
    with tempfile.TemporaryDirectory() as td:
        db100 = base.bootstrap(100000, 20240502, True)
        db100.write_packed(os.path.join(td, "db100.satsdb"))
        db100.write_ascii(os.path.join(td, "db100.ascii"))
        q = qs.select([qs.find("D2PHLB1")])
        q.write_ascii(os.path.join(td, "q.ascii"))
        qtext = Path(td, "q.ascii").read_text()
        for dbfile in ("db100.satsdb", "db100.ascii"):
            Path(td, "in_" + dbfile).write_text("%s\nT T F\n%s" % (dbfile, qtext))
        db15 = base.bootstrap(14297, 20240501, True)
        db15.write_packed(os.path.join(td, "db15.satsdb"))
        rng = np.random.default_rng(200)
        ids = [db15.name(int(i)) for i in rng.choice(len(db15), 200, replace=False)]
        Path(td, "ids200").write_text("\n".join(ids) + "\n")
        run_cli(["-r", 128, "-g", 1], os.path.join(td, "in_db100.satsdb"), td)          # warm the driver / page cache
        md5 = {}
        for n in gpus:
            if n > have and n > 1 and have > 0 and "--oversubscribe" not in sys.argv:
                continue
            for what, argv, stdin, pairs in (
                    ("configs[4]: D2PHLB1 vs 100k db (packed cache), R=128", ["-r", 128, "-g", n], "in_db100.satsdb", 100000),
                    ("configs[4]: the same from the ASCII db (148 MB of text)", ["-r", 128, "-g", n], "in_db100.ascii", 100000),
                    ("configs[2]: 200 queries (-q) vs 14 297 db, R=128", ["-q", "db15.satsdb", "-r", 128, "-g", n], "ids200", 200 * 14297)):
                r = run_cli(argv, os.path.join(td, stdin), td)
                key = what.split(":")[0] + stdin
                md5.setdefault(key, r["stdout_md5"])
                r.update({"workload": what, "gpus": n, "gpus_present": have, "pairs": pairs,
                          "structures_per_s_search": pairs / (r["search_ms"] / 1e3),
                          "structures_per_s_wall": pairs / r["wall_s"],
                          "stdout_identical_to_first_run": r["stdout_md5"] == md5[key]})
                print(json.dumps(r), flush=True)


if __name__ == "__main__":
    main()
