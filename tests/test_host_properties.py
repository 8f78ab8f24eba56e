"""Property tests (hypothesis) of the host half of libsats: any structure list survives ASCII and packed round trips,
and the result formatter/reader are inverse to each other."""
import numpy as np
from hypothesis import given, settings, strategies as st

import cuda_satabsearch_b200 as S


@st.composite
def structures(draw):
    count = draw(st.integers(1, 6))
    out = []
    for k in range(count):
        n = draw(st.sampled_from([1, 2, 3, 5, 9, 16, 33, 70, 111]))
        rng = np.random.default_rng(draw(st.integers(0, 2 ** 32 - 1)))
        tab = np.zeros((n, n), np.uint8)
        dm = np.zeros((n, n), np.float32)
        for i in range(n):
            t = int(rng.integers(0, 4))
            tab[i, i] = t
            dm[i, i] = t
            for j in range(i):
                tab[i, j] = tab[j, i] = (int(rng.integers(0, 5)) << 4) | int(rng.integers(0, 5))
                dm[i, j] = dm[j, i] = np.float32(int(rng.integers(0, 99999)) / 1000.0)     # what "%6.3f" can carry
        out.append(("s%04d%c" % (draw(st.integers(0, 9999)), "abc"[k % 3]), tab, dm))
    return out


@settings(max_examples=25, deadline=None)
@given(structures())
def test_ascii_and_packed_roundtrip(tmp_path_factory, items):
    d = tmp_path_factory.mktemp("rt")
    db = S.Database.from_structures([x[0] for x in items], [x[1] for x in items], [x[2] for x in items])
    db.write_ascii(d / "x.ascii")
    db.write_packed(d / "x.satsdb")
    for back in (S.Database.read_ascii(d / "x.ascii"), S.Database.read_packed(d / "x.satsdb")):
        assert len(back) == len(items)
        for k, (name, tab, dm) in enumerate(items):
            t, dd = back.get(k)
            assert back.name(k) == name and np.array_equal(t, tab) and np.array_equal(dd, dm)


@settings(max_examples=25, deadline=None)
@given(st.lists(st.integers(-300, 400), min_size=1, max_size=40), st.integers(1, 111), st.booleans())
def test_formatter_and_reader_are_inverse(scores, qn, lorder):
    n = len(scores)
    names = ["e%03d" % k for k in range(n)]
    tabs = [np.zeros((1 + k % 7, 1 + k % 7), np.uint8) for k in range(n)]
    dms = [np.zeros((1 + k % 7, 1 + k % 7), np.float32) for k in range(n)]
    db = S.Database.from_structures(names, tabs, dms)
    text = db.format_block("QUERYID", qn, "some/db.ascii", lorder, False, np.array(scores, np.int32))
    blk = S.parse_results(text)
    assert len(blk) == 1 and blk[0]["query"] == "QUERYID" and blk[0]["lorder"] == lorder
    assert blk[0]["names"] == names and blk[0]["scores"].tolist() == scores
    for k in range(n):
        assert abs(blk[0]["norm2"][k] - S.norm2(scores[k], qn, 1 + k % 7)) <= 1e-5 * max(1.0, abs(blk[0]["norm2"][k]))


def test_ascii_writer_formats_distances_like_printf(tmp_path):
    """The writer's hand-rolled "%6.3f " must equal printf's for every float: exact ties at the third decimal (round half
    to even), the 99.9995 boundary, values that need more than six characters, negatives, -0, NaN (-> 0.000), random bit
    patterns."""
    rng = np.random.default_rng(99)
    vals = [0.0, -0.0, 0.0005, 0.0015, 0.0625, 0.1875, 0.3125, 2.5e-4, 9.9995, 9.99949, 99.9994, 99.9995, 99.99951, 100.0, 123.456,
            1e6, -1.0, -0.0004, -12.3456, float("nan"), 1e-30, 5e-4, 1.0005, 31.4155, 31.4165, 64.0625, 7.8125e-3]
    vals += [k / 1000.0 + 0.0005 for k in range(0, 4000, 37)]                    # near-ties as doubles -> float
    vals += [float(x) for x in rng.uniform(0, 100, 3000)]
    vals += [float(np.frombuffer(np.uint32(b).tobytes(), np.float32)[0]) for b in rng.integers(0x30000000, 0x42C80000, 3000)]
    vals = np.array(vals, np.float32)
    n = 111
    per = n * (n - 1) // 2
    structs = []
    for s in range(0, len(vals), per):
        chunk = vals[s:s + per]
        dm = np.zeros((n, n), np.float32)
        iu = np.tril_indices(n, -1)
        dm[iu[0][:len(chunk)], iu[1][:len(chunk)]] = chunk
        dm = dm + dm.T
        structs.append(("w%05d" % s, np.zeros((n, n), np.uint8), dm))
    db = S.Database.from_structures([x[0] for x in structs], [x[1] for x in structs], [x[2] for x in structs])
    db.write_ascii(tmp_path / "w.ascii")
    lines = (tmp_path / "w.ascii").read_text().split("\n")
    pos = 0
    for name, _, dm in structs:
        while not lines[pos].strip():
            pos += 1
        assert lines[pos].split()[0] == name
        pos += 1 + n                                                              # header + tableau rows
        for i in range(n):
            want = "".join("%6.3f " % (0.0 if np.isnan(dm[i, j]) else float(dm[i, j])) for j in range(i + 1))
            assert lines[pos + i] == want, (name, i)
        pos += n


# ---- the kernel's integer cut-offs against the oracle's float / double formulas (reference kernel.cu:1042, :1166, :624)
def _around(cuts, rng, extra=2000):
    pts = set()
    for c in np.asarray(cuts, np.uint64).ravel().tolist():
        for d in (-2, -1, 0, 1, 2):
            if 0 <= c + d <= 0xffffffff:
                pts.add(int(c + d))
    pts |= {0, 1, 0xffffffff, 0xfffffffe, 0x7fffffc0, 0x7fffffbf, 0x80000000}
    pts |= set(int(v) for v in rng.integers(0, 2**32, extra, dtype=np.uint64))
    return sorted(pts)


def test_sse_pick_boundaries_reproduce_the_reference_pick(oracle):
    """pick_index() in the kernel = umulhi(x, n) - (x < cut[umulhi(x, n)]); it must equal the reference's
    (int)((unit(x) - 1.1e-7) * n) for every x: checked at and around every boundary and on random draws, for every n."""
    import cuda_satabsearch_b200 as S
    f = oracle.lib.sats_oracle_pick_from_bits
    rng = np.random.default_rng(11)
    for n in range(1, 112):
        cut = S.pick_boundaries(n)
        assert cut[0] == 0 and (np.diff(cut.astype(np.int64)) > 0).all()
        for x in _around(cut, rng, 300 if n % 10 else 3000):
            s0 = (x * n) >> 32
            assert s0 - (1 if x < int(cut[s0]) else 0) == f(x, n), (n, x)


def test_accept_cutoffs_reproduce_the_reference_metropolis_test(oracle):
    import cuda_satabsearch_b200 as S
    cut, temps = S.accept_cutoffs()
    f = oracle.lib.sats_oracle_accept_from_bits
    thr = oracle.lib.sats_oracle_accept_threshold
    rng = np.random.default_rng(12)
    t = np.float32(10.0)
    for m in range(100):
        assert temps[m] == t
        t = np.float32(t * np.float32(0.95))
    assert (cut[:, 0] == cut[0, 0]).all() and cut[0, 0] == 0xffffff80          # d == 0: unit(x) < 1.0f
    assert (np.diff(cut.astype(np.int64), axis=1) <= 0).all()                    # a larger loss is never easier to accept
    for m in (0, 1, 17, 50, 98, 99):
        for nd in list(range(0, 40)) + [77, 150, 229]:
            c = int(cut[m, nd])
            for x in {max(c - 1, 0), c, min(c + 1, 0xffffffff), 0, 0xffffffff, int(rng.integers(0, 2**32))}:
                assert (x < c) == bool(f(m, -nd, x)), (m, nd, x, c, thr(m, -nd))
    # beyond the table: expf(-d / T) <= 2^-33 for d >= 229 even at T = 10, so such moves never pass
    assert not f(0, -230, 0) and int(cut[0, 229]) == 0


def test_seed_cutoff(oracle):
    import cuda_satabsearch_b200 as S
    c = S.seed_cutoff()
    f = oracle.lib.sats_oracle_seed_attempt_from_bits
    assert f(c - 1) == 1 and f(c) == 0 and f(0) == 1 and f(0xffffffff) == 0
