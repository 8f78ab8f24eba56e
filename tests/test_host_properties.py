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
