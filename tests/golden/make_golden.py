#!/usr/bin/env python
"""Regenerate tests/golden/ from the reference checkout (/root/reference) and the reference binary
built by oracle/build_ref.sh (oracle/_ref/cudaSaTabsearch_ref[_md32]).

Runs only in the build container (the GPU box has no /root/reference).  What it writes:

  *.satsdb       the reference's fixture databases / query structures re-encoded in this project's packed
                 binary format (tests/_refio.py: write_packed) -- inputs for every parity test.
  golden.json    for each case: the exact command, md5 of the reference's stdout, the per-entry scores
                 (and SSE maps when LSOLN=T) parsed from that stdout; the md5 of each ASCII fixture (to pin
                 our ASCII writer byte for byte); and the reference's captured 2013 job output
                 old/nvcc_src_cuda5/cpu_cudaSaTabsearch.o1462445 reduced to (md5, names, scores).

Usage:  python tests/golden/make_golden.py
"""
import hashlib
import json
import os
import subprocess
import sys
import tempfile
from pathlib import Path

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent))
from _refio import parse_ascii_db, parse_query_input, write_packed  # noqa: E402

REF = Path(os.environ.get("SATS_REFERENCE_DIR", "/root/reference"))
SRC = REF / "nvcc_src_current"
BIN = HERE.parent.parent / "oracle" / "_ref" / "cudaSaTabsearch_ref"
BIN32 = HERE.parent.parent / "oracle" / "_ref" / "cudaSaTabsearch_ref_md32"

DBS = {
    "small586": "tableauxdistmatrixdb.small.ascii",
    "test1": "tableauxdistmatrixdb.test.ascii",
    "test2": "tableauxdistmatrixdb.test2.ascii",
    "d1qlpa": "d1qlpa_.ascii",
    "d2pq6a1": "d2pq6a1.ascii",
}
QUERY_FILES = ["d1ubia_.input", "d2phlb1.input", "d1ae6h1.input", "d1twfa_.input", "1qlp_sheetbc.input",
               "2qp2-1.input"]

# (case id, query input file, db key, lorder, lsoln, restarts, binary, pool threshold)
CASES = [
    ("d1ubia_small_r128", "d1ubia_.input", "small586", True, False, 128, BIN, 96),       # BASELINE config 1
    ("d1ubia_test1_default", "d1ubia_.input", "test1", True, True, 128, BIN, 96),        # known answer 54 / identity
    ("d1ae6h1_test2_r128", "d1ae6h1.input", "test2", True, False, 128, BIN, 96),         # known answer 92
    ("d2phlb1_small_r128", "d2phlb1.input", "small586", True, False, 128, BIN, 96),
    ("multiquery_small_r128", "multiquery.input", "small586", True, False, 128, BIN, 96),
    ("sheetbc_d1qlpa_TTT_r1024", "1qlp_sheetbc.input", "d1qlpa", True, True, 1024, BIN, 96),
    ("sheetbc_d1qlpa_TFT_r1024", "1qlp_sheetbc.input", "d1qlpa", False, True, 1024, BIN, 96),
    ("sheetbc_small_TFT_r128", "1qlp_sheetbc.input", "small586", False, True, 128, BIN, 96),
    ("d2phlb1_d2pq6a1_TTT_r128", "d2phlb1.input3", "d2pq6a1", True, True, 128, BIN, 96),
    ("d2phlb1_small_r128_md32", "d2phlb1.input", "small586", True, False, 128, BIN32, 32),  # small+large pools
]


def md5(b: bytes) -> str:
    return hashlib.md5(b).hexdigest()


def parse_stdout(text: str):
    """-> list of blocks {query, rows:[name...], scores:[...], maps:[[ [k,j],...], ...]}"""
    blocks, cur = [], None
    for line in text.split("\n"):
        if line.startswith("# cudaSaTabsearch"):
            cur = {"query": None, "names": [], "scores": [], "maps": []}
            blocks.append(cur)
        elif line.startswith("# QUERY ID"):
            cur["query"] = line.split("=")[1].strip()
        elif line.startswith("#") or not line.strip():
            continue
        else:
            f = line.split()
            if len(f) == 5:
                cur["names"].append(f[0]); cur["scores"].append(int(f[1])); cur["maps"].append([])
            elif len(f) == 2:
                cur["maps"][-1].append([int(f[0]), int(f[1])])
    return blocks


def run_case(binary, inp, dbname, lorder, lsoln, restarts):
    text = (SRC / inp).read_text().split("\n")
    text[0] = dbname
    text[1] = "T %s %s" % ("T" if lorder else "F", "T" if lsoln else "F")
    p = subprocess.run([str(binary), "-c", "-r", str(restarts)], input="\n".join(text).encode(),
                       cwd=str(SRC), capture_output=True, check=True)
    return p.stdout


def main():
    out = {"ascii_md5": {}, "cases": {}, "captured": {}}
    for key, fname in DBS.items():
        ents = parse_ascii_db(SRC / fname)
        write_packed(HERE / (key + ".satsdb"), ents)
        raw = (SRC / fname).read_bytes()
        out["ascii_md5"][key] = {"file": fname, "md5": md5(raw), "bytes": len(raw), "entries": len(ents)}
    queries, seen = [], set()
    for qf in QUERY_FILES + ["multiquery.input"]:
        for q in parse_query_input((SRC / qf).read_text())[4]:
            if q.name not in seen:
                seen.add(q.name); queries.append(q)
    write_packed(HERE / "queries.satsdb", queries)
    out["queries"] = [q.name for q in queries]

    for cid, inp, dbkey, lorder, lsoln, restarts, binary, thr in CASES:
        stdout = run_case(binary, inp, DBS[dbkey], lorder, lsoln, restarts)
        blocks = parse_stdout(stdout.decode())
        if not lsoln:
            for b in blocks:
                b.pop("maps")
        out["cases"][cid] = {
            "command": "%s -c -r %d < %s  (line 1 -> %s, line 2 -> T %s %s)" % (
                binary.name, restarts, inp, DBS[dbkey], "T" if lorder else "F", "T" if lsoln else "F"),
            "input": inp, "db": dbkey, "dbfile": DBS[dbkey], "lorder": lorder, "lsoln": lsoln,
            "restarts": restarts, "pool_threshold": thr, "stdout_md5": md5(stdout), "blocks": blocks,
        }
        print(cid, md5(stdout), sum(len(b["scores"]) for b in blocks), "rows")

    cap = REF / "old" / "nvcc_src_cuda5" / "cpu_cudaSaTabsearch.o1462445"
    raw = cap.read_bytes()
    blocks = parse_stdout(raw.decode())
    for b in blocks:
        b.pop("maps")
    out["captured"]["cpu_2013_d2phlb1_r4096"] = {
        "source": "old/nvcc_src_cuda5/cpu_cudaSaTabsearch.o1462445 (2013 CPU job, MAXDIM_GPU 32, -r4096 < d2phlb1.input)",
        "input": "d2phlb1.input", "db": "small586", "dbfile": DBS["small586"], "lorder": True, "lsoln": False,
        "restarts": 4096, "pool_threshold": 32, "stdout_md5": md5(raw), "blocks": blocks,
    }
    (HERE / "golden.json").write_text(json.dumps(out, separators=(",", ":")) + "\n")
    print("wrote", HERE / "golden.json")


if __name__ == "__main__":
    main()
