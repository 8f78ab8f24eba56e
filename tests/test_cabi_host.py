"""CPU-side tests of libsats.so: the library loads, exports every symbol include/sats.h declares, and its host
logic (parser, packed cache, ASCII writer, statistics, result formatter, partitioner, stdin grammars) behaves like
the reference's host code.  No search is run here (that needs a GPU and must fail loudly without one)."""
import hashlib
import re
import subprocess
import sys

import numpy as np
import pytest

import cuda_satabsearch_b200 as S
from _refio import GOLDEN, REPO, format_entry, read_packed, stats


def test_library_exports_every_declared_symbol():
    hdr = (REPO / "include" / "sats.h").read_text()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = set(re.findall(r"\b(sats_[a-z0-9_]+)\s*\(", hdr))
    names |= set(re.findall(r"extern const double (sats_[a-z_]+);", hdr))
    assert len(names) >= 35
    out = subprocess.run(["nm", "-D", "--defined-only", str(S.LIB_PATH)], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r"\b[TDRB] (sats_[a-z0-9_]+)", out))
    assert names - exported == set(), names - exported
    S.lib()          # binds every function; AttributeError if one is missing


def test_no_cpu_fallback_without_gpu():
    if S.device_count() > 0:
        pytest.skip("GPU present")
    db = S.Database.read_packed(GOLDEN / "test1.satsdb")
    with pytest.raises(S.SatsError, match="no CUDA device"):
        S.Searcher(db)


def test_product_never_touches_the_oracle():
    for p in list((REPO / "cuda_satabsearch_b200").rglob("*")) + [REPO / "include" / "sats.h"]:
        if p.is_file() and p.suffix in (".py", ".c", ".cpp", ".cu", ".cuh", ".h") or p.name == "Makefile":
            txt = p.read_text(errors="ignore")
            assert "sats_oracle" not in txt and "oracle/" not in txt.replace("the oracle", ""), p


@pytest.mark.parametrize("key", ["small586", "test1", "test2", "d1qlpa", "d2pq6a1", "queries"])
def test_packed_reader_matches_independent_reader(key):
    db = S.Database.read_packed(GOLDEN / (key + ".satsdb"))
    ref = read_packed(GOLDEN / (key + ".satsdb"))
    assert len(db) == len(ref)
    step = max(1, len(ref) // 60)
    for i in list(range(0, len(ref), step)) + [len(ref) - 1]:
        t, d = db.get(i)
        assert db.name(i) == ref[i].name and db.order(i) == ref[i].n
        assert np.array_equal(t, ref[i].tab) and np.array_equal(d, ref[i].dmat)


def test_ascii_writer_is_byte_identical_to_reference_files(tmp_path, golden):
    for key, meta in golden["ascii_md5"].items():
        db = S.Database.read_packed(GOLDEN / (key + ".satsdb"))
        out = tmp_path / (key + ".ascii")
        db.write_ascii(out)
        raw = out.read_bytes()
        if len(raw) + 1 == meta["bytes"]:
            raw += b"\n"
        assert hashlib.md5(raw).hexdigest() == meta["md5"], key


def test_ascii_parse_roundtrip_and_packed_roundtrip(tmp_path):
    ref = read_packed(GOLDEN / "small586.satsdb")
    text = "\n".join(format_entry(s) for s in ref)
    db = S.Database.parse_ascii(text)
    assert len(db) == 586
    for i in (0, 1, 17, 300, 585):
        t, d = db.get(i)
        assert np.array_equal(t, ref[i].tab) and np.array_equal(d, ref[i].dmat) and db.name(i) == ref[i].name
    db.write_packed(tmp_path / "x.satsdb")
    assert (tmp_path / "x.satsdb").read_bytes() == (GOLDEN / "small586.satsdb").read_bytes()


def test_parser_edge_cases():
    assert len(S.Database.parse_ascii("")) == 0                       # empty db
    one = "A          1\ne  \n 0.000 \n"
    db = S.Database.parse_ascii(one)
    assert len(db) == 1 and db.order(0) == 1 and db.get(0)[0][0, 0] == 0
    with pytest.raises(S.SatsError, match="invalid tableaux code"):
        S.Database.parse_ascii("B          2\ne  \nXX e  \n 0.000 \n 1.000  0.000 \n")
    with pytest.raises(S.SatsError, match="Bad helix type"):
        S.Database.parse_ascii("B          1\nxz \n 0.000 \n")
    with pytest.raises(S.SatsError, match="truncated"):
        S.Database.parse_ascii("B          2\ne  \nPE e  \n 0.000 \n")
    # structures larger than MAXDIM are skipped with a warning, the rest is kept (parsetableaux.c:457-465)
    n = 112
    big = "BIG      %4d\n" % n + "".join("PE " * i + "e  \n" for i in range(n)) + "".join(" 1.000 " * (i + 1) + "\n" for i in range(n))
    db = S.Database.parse_ascii(big + "\n" + one)
    assert len(db) == 1 and db.name(0) == "A"
    # SURVEY 8(f4), opt-in: the same text with max_order=128 keeps the big structure (and 129 is refused)
    db = S.Database.parse_ascii(big + "\n" + one, max_order=S.MAXDIM_EXT)
    assert len(db) == 2 and db.name(0) == "BIG" and db.order(0) == 112 and db.get(0)[1][111, 0] == np.float32(1.0)
    with pytest.raises(S.SatsError, match="max_order"):
        S.Database.parse_ascii(one, max_order=S.MAXDIM_EXT + 1)
    # the 5x5 code alphabet incl. '?' (parsetableaux.c:92-138)
    q = "C          2\nxg \n?? xi \n 3.000 \n 9.999  2.000 \n"
    t, d = S.Database.parse_ascii(q).get(0)
    assert t.tolist() == [[3, 0x44], [0x44, 2]] and d[1, 0] == np.float32(9.999)


def test_input_and_idlist_grammar():
    ref = read_packed(GOLDEN / "queries.satsdb")
    body = "\n".join(format_entry(s) for s in ref[:3])
    dbf, lt, lo, ls, qs = S.parse_input("some/db.ascii\nT F T\n" + body)
    assert (dbf, lt, lo, ls) == ("some/db.ascii", True, False, True)
    assert qs.names() == [s.name for s in ref[:3]]
    with pytest.raises(S.SatsError, match="no query structures"):
        S.parse_input("db\nT T F\n")
    # ids are cut to 7 characters like cudaSaTabsearch.cu:657
    assert S.parse_idlist("d1ubia_\nd2phlb1xyz\n\nabc\n") == ["d1ubia_", "d2phlb1", "abc"]
    db = S.Database.read_packed(GOLDEN / "small586.satsdb")
    assert db.find("D1KCUL1") == 0                                    # strcasecmp, cudaSaTabsearch.cu:747
    with pytest.raises(S.SatsError, match="not found"):
        db.find("nosuch")


def test_statistics_and_formatter_reproduce_reference_stdout(golden, oracle):
    """Scores from the golden reference run -> sats_format_block must re-create the reference's stdout."""
    for cid in ("d1ubia_small_r128", "d2phlb1_small_r128", "sheetbc_small_TFT_r128", "d1ubia_test1_default"):
        case = golden["cases"][cid]
        db = S.Database.read_packed(GOLDEN / (case["db"] + ".satsdb"))
        blk = case["blocks"][0]
        qs = S.Database.read_packed(GOLDEN / "queries.satsdb")
        qn = qs.order(qs.find(blk["query"]))
        scores = np.array(blk["scores"], np.int32)
        maps = None
        if case["lsoln"]:
            maps = np.full((len(db), S.MAP_STRIDE), -1, np.int32)
            for e, pairs in enumerate(blk["maps"]):
                for k, j in pairs:
                    maps[e, k - 1] = j - 1
        text = db.format_block(blk["query"], qn, case["dbfile"], case["lorder"], case["lsoln"], scores, maps)
        assert hashlib.md5(text.encode()).hexdigest() == case["stdout_md5"], cid
    for score, n1, n2 in [(54, 8, 8), (-3, 19, 40), (0, 9, 1), (7, 101, 67)]:
        n2s, z, p = stats(score, n1, n2)
        assert S.norm2(score, n1, n2) == n2s and S.z_gumbel(int(n2s)) == z and S.pv_gumbel(z) == p


def test_partition_is_balanced_and_complete():
    db = S.Database.read_packed(GOLDEN / "small586.satsdb")
    orders = db.orders()
    for n in (1, 2, 3, 8):
        own = db.partition(n)
        assert own.min() == 0 and own.max() == n - 1
        cost = 13.0 + orders
        loads = np.array([cost[own == r].sum() for r in range(n)])
        assert loads.max() - loads.min() <= cost.max() + 1e-9          # LPT bound
    big = db.bootstrap(20000, 20240501, True)
    o = big.orders()
    assert len(big) == 20000 and np.all(np.diff(o) >= 0) and big.name(0).startswith("s")
    assert abs(o.mean() - orders.mean()) < 0.5
    again = db.bootstrap(20000, 20240501, True)
    assert np.array_equal(again.orders(), o)                            # deterministic


def test_select_and_from_structures_roundtrip():
    ref = read_packed(GOLDEN / "small586.satsdb")[:20]
    db = S.Database.from_structures([s.name for s in ref], [s.tab for s in ref], [s.dmat for s in ref])
    sub = db.select([5, 0, 19])
    assert sub.names() == [ref[5].name, ref[0].name, ref[19].name]
    assert np.array_equal(sub.get(2)[1], ref[19].dmat)


def test_results_reader_and_auc(golden):
    """the five-column output reader round-trips what the formatter prints, and sats_roc_auc equals
    the Mann-Whitney statistic."""
    case = golden["cases"]["d1ubia_test1_default"]          # LSOLN=T: rows + map pairs
    db = S.Database.read_packed(GOLDEN / "test1.satsdb")
    blk = case["blocks"][0]
    maps = np.full((1, S.MAP_STRIDE), -1, np.int32)
    for k, j in blk["maps"][0]:
        maps[0, k - 1] = j - 1
    text = db.format_block("D1UBIA_", 8, case["dbfile"], True, True, np.array(blk["scores"], np.int32), maps)
    case2 = golden["cases"]["d2phlb1_small_r128"]
    db2 = S.Database.read_packed(GOLDEN / "small586.satsdb")
    text += db2.format_block("D2PHLB1", 19, case2["dbfile"], True, False, np.array(case2["blocks"][0]["scores"], np.int32))
    out = S.parse_results(text)
    assert [b["query"] for b in out] == ["D1UBIA_", "D2PHLB1"]
    assert out[0]["lsoln"] and not out[1]["lsoln"] and out[0]["dbfile"] == case["dbfile"]
    assert out[0]["names"] == ["d1ndda_"] and out[0]["scores"].tolist() == [54]
    assert out[0]["maps"][0].tolist() == [[k, k] for k in range(1, 9)]
    assert out[1]["names"] == case2["blocks"][0]["names"] and out[1]["scores"].tolist() == case2["blocks"][0]["scores"]
    n2s, z, p = stats(54, 8, 8)
    assert abs(out[0]["norm2"][0] - n2s) < 1e-5 and abs(out[0]["z"][0] - z) < 1e-4
    with pytest.raises(S.SatsError, match="not understood"):
        S.parse_results("name 1 2\n")
    rng = np.random.default_rng(3)
    score = rng.integers(0, 12, 400).astype(float)
    y = rng.random(400) < 0.2
    pos, neg = score[y], score[~y]
    want = ((pos[:, None] > neg[None, :]).sum() + 0.5 * (pos[:, None] == neg[None, :]).sum()) / (len(pos) * len(neg))
    assert abs(S.roc_auc(score, y) - want) < 1e-12
    assert np.isnan(S.roc_auc(score, np.zeros(400, bool)))


def test_c_consumer_links_and_runs(tmp_path):
    """include/sats.h must be plain C and libsats must link from a C program (the reference's host code is C)."""
    exe = tmp_path / "cabi_smoke"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", str(REPO / "include"),
                    str(REPO / "tests" / "cabi_smoke.c"), "-o", str(exe), "-L", str(S.LIB_PATH.parent), "-lsats",
                    "-Wl,-rpath," + str(S.LIB_PATH.parent), "-lm"], check=True, capture_output=True)
    p = subprocess.run([str(exe), str(GOLDEN / "small586.satsdb")], capture_output=True, text=True, timeout=120)
    assert p.returncode == 0, (p.returncode, p.stdout, p.stderr)
    first = p.stdout.split("\n")[0].split()
    assert first[:2] == ["586", "67"] and abs(float(first[2]) - 6.75) < 1e-9 and abs(float(first[3]) - 11.7853) < 1e-3
    if S.device_count() == 0:
        assert "no CUDA device" in p.stdout


def test_score_threshold_is_the_first_score_that_passes():
    """sats_score_threshold (the cut sats_search_hits applies on the device) against the printer's own z-score formula."""
    rng = np.random.default_rng(5)
    for _ in range(300):
        n1, n2 = int(rng.integers(1, 112)), int(rng.integers(1, 129))
        z = float(rng.uniform(-3.0, 40.0))
        t = S.score_threshold(z, n1, n2)
        if t == 2**31 - 1:
            assert stats(16384, n1, n2)[1] < z
            continue
        assert stats(t, n1, n2)[1] >= z and stats(t - 1, n1, n2)[1] < z, (n1, n2, z, t)
    assert S.score_threshold(-1e9, 5, 7) == -16384 and S.score_threshold(float("nan"), 5, 7) == 2**31 - 1


def test_fast_distance_parse_equals_strtof_for_every_canonical_value():
    """The parser reads the writer's "%6.3f" fields through a table of strtof results; a trailing zero pushes the same
    values through strtof itself.  All 100 000 of them must agree bit for bit."""
    fast = "".join("s%05d     1\ne  \n%6.3f \n\n" % (k % 100000, k / 1000.0) for k in range(100000))
    slow = "".join("s%05d     1\ne  \n%d.%03d0 \n\n" % (k % 100000, k // 1000, k % 1000) for k in range(100000))
    a, b = S.Database.parse_ascii(fast), S.Database.parse_ascii(slow)
    assert len(a) == len(b) == 100000
    va = np.array([a.get(i)[1][0, 0] for i in range(100000)], np.float32)
    vb = np.array([b.get(i)[1][0, 0] for i in range(100000)], np.float32)
    assert np.array_equal(va.view(np.uint32), vb.view(np.uint32))
    assert abs(float(va[-1]) - 99.999) < 1e-5 and (np.diff(va) > 0).all()


def test_constructors_reject_codes_outside_the_alphabet(tmp_path):
    """Only the ASCII parser used to validate codes; arrays and the packed cache now apply the same alphabet (the kernel
    indexes shared-memory tables with the SSE type and the letters)."""
    tab = np.array([[0, 0x11], [0x11, 1]], np.uint8)
    dm = np.zeros((2, 2), np.float32)
    S.Database.from_structures(["ok"], [tab], [dm])
    for bad in ([[4, 0x11], [0x11, 1]], [[0, 0x51], [0x51, 1]], [[0, 0x15], [0x15, 1]], [[0, 0xff], [0xff, 3]]):
        with pytest.raises(S.SatsError, match="invalid"):
            S.Database.from_structures(["bad"], [np.array(bad, np.uint8)], [dm])
    good = tmp_path / "good.satsdb"
    S.Database.from_structures(["ok"], [tab], [dm]).write_packed(good)
    raw = bytearray(good.read_bytes())
    assert len(S.Database.read_packed(good)) == 1
    tab_at = 24 + 16                       # header, then order[1] + name[9] padded to 16
    raw[tab_at] = 9                        # SSE type 9 on the diagonal
    bad = tmp_path / "bad.satsdb"
    bad.write_bytes(bytes(raw))
    with pytest.raises(S.SatsError, match="invalid SSE type"):
        S.Database.read_packed(bad)


def test_packed_reader_survives_corrupt_headers(tmp_path):
    src = GOLDEN / "test1.satsdb"
    raw = bytearray(src.read_bytes())
    for cells in (0x3333333333333334, 0xffffffffffffffff, 1 << 40):        # wrap-around and absurd sizes
        r = bytearray(raw)
        r[16:24] = int(cells).to_bytes(8, "little")
        p = tmp_path / "c.satsdb"
        p.write_bytes(bytes(r))
        with pytest.raises(S.SatsError, match="truncated|mismatch"):
            S.Database.read_packed(p)
    r = bytearray(raw)
    r[8:12] = (0xffffffff).to_bytes(4, "little")
    (tmp_path / "n.satsdb").write_bytes(bytes(r))
    with pytest.raises(S.SatsError, match="truncated"):
        S.Database.read_packed(tmp_path / "n.satsdb")
    r = bytearray(raw)
    r[24:28] = (5000).to_bytes(4, "little")                                   # an order the cell count cannot hold
    (tmp_path / "o.satsdb").write_bytes(bytes(r))
    with pytest.raises(S.SatsError, match="order|mismatch"):
        S.Database.read_packed(tmp_path / "o.satsdb")


def test_shim_checks_caller_buffers():
    from cuda_satabsearch_b200 import _out_array
    ok = np.zeros((2, 5), np.int32)
    assert _out_array(ok, (2, 5), "scores") is ok
    for bad in (np.zeros((2, 5), np.float64), np.zeros((2, 4), np.int32), np.zeros((5, 2), np.int32).T, np.zeros((2, 10), np.int32)[:, ::2]):
        with pytest.raises(S.SatsError, match="scores must be"):
            _out_array(bad, (2, 5), "scores")


def test_result_printer_tail_cache_is_transparent():
    """sats_format_block memoizes the numeric columns per (score, structure order); with far more distinct pairs than the
    cache holds (and negative scores) every row must still be what printf makes of it."""
    db = S.Database.read_packed(GOLDEN / "small586.satsdb").bootstrap(30000, 3, False)
    rng = np.random.default_rng(9)
    sc = rng.integers(-3000, 12000, len(db)).astype(np.int32)
    txt = db.format_block("QUERYID", 23, "some/db.ascii", True, False, sc)
    rows = txt.split("\n")
    assert rows[1] == "# QUERY ID = QUERYID " and len(rows) == 3 + len(db) + 1
    orders = db.orders()
    for k in list(range(0, len(db), 7)) + [len(db) - 1]:
        n2s, z, p = stats(int(sc[k]), 23, int(orders[k]))
        assert rows[3 + k] == "%-8s %d %g %g %g" % (db.name(k), int(sc[k]), n2s, z, p), k


def test_parallel_ascii_parse_equals_serial(tmp_path):
    """Database texts above 4 MB are cut at blank lines and parsed on several threads; the result -- structures, their order,
    the oversize warnings and their order -- must be what the serial parse gives, and anything irregular (here: a piece
    that fails) must fall back to the serial parse and its error message."""
    import os
    base = S.Database.read_packed(GOLDEN / "small586.satsdb")
    db = base.bootstrap(9000, 77, False)
    path = tmp_path / "db.ascii"
    db.write_ascii(path)
    text = path.read_text()
    assert len(text) > 8 << 20
    # plant oversize entries (order 112 > 111) at a few places: they are skipped with warnings, in file order
    big = "toobig1  112\n" + "\n".join(" ".join(["e "] * (i + 1)) + " " for i in range(112)) + "\n" + \
          "\n".join(" ".join(["%6.3f" % 1.0] * (i + 1)) + " " for i in range(112)) + "\n"
    entries = text.split("\n\n")
    for k, at in enumerate((10, 4000, 8990)):
        entries.insert(at + k, big.replace("toobig1", "toobig%d" % k).rstrip("\n"))
    text = "\n\n".join(entries)
    (tmp_path / "w.ascii").write_text(text)
    code = ("import sys; sys.path.insert(0, %r); import cuda_satabsearch_b200 as S; d = S.Database.read_ascii(%r); "
            "d.write_packed(%r)")
    outs = {}
    for thr in ("1", "6"):
        env = dict(os.environ, SATS_PARSE_THREADS=thr)
        p = subprocess.run([sys.executable, "-c", code % (str(REPO), str(tmp_path / "w.ascii"), str(tmp_path / ("p%s.satsdb" % thr)))],
                           capture_output=True, text=True, env=env, timeout=300)
        assert p.returncode == 0, p.stderr[-500:]
        outs[thr] = (hashlib.md5((tmp_path / ("p%s.satsdb" % thr)).read_bytes()).hexdigest(), p.stderr)
    assert outs["1"] == outs["6"]
    assert outs["1"][1].count("is too large") == 6 and "skipped 3 database tableaux" in outs["1"][1]
    assert outs["1"][1].index("toobig0") < outs["1"][1].index("toobig1") < outs["1"][1].index("toobig2")
    # a broken entry in the middle: same error either way
    bad = text.replace("toobig1  112", "toobig1  112x", 1)
    msgs = []
    for thr in ("1", "6"):
        os.environ["SATS_PARSE_THREADS"] = thr
        try:
            n = len(S.Database.parse_ascii(bad))
            msgs.append("parsed %d" % n)
        except S.SatsError as exc:
            msgs.append(str(exc))
    os.environ.pop("SATS_PARSE_THREADS", None)
    assert msgs[0] == msgs[1], msgs
