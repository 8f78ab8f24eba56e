"""SURVEY 8(f3), query construction: the C ABI's angle -> tableau code, interaxial angle and structure builder against
golden vectors produced by executing the reference's own pure functions (tests/golden/make_f3_golden.py:
scripts/pttableau.py:434-469, scripts/ptnode.py:752-880, scripts/geometry.py:18-79)."""
import json
import math

import numpy as np
import pytest

import cuda_satabsearch_b200 as S
from _refio import GOLDEN


@pytest.fixture(scope="module")
def f3():
    return json.loads((GOLDEN / "f3_golden.json").read_text())


def test_angle_to_tabcode_matches_the_reference_on_every_edge(f3):
    """Exact: every interval edge (+- pi/4, pi/2, 3pi/4, pi, 0) with its floating-point neighbours, a dense sweep, random
    angles; -pi, anything outside (-pi, pi] and NaN raise (ValueError in the reference, SATS_ERR_ARG here)."""
    seen = set()
    for rec in f3["codes"]:
        om = float("nan") if rec["omega"] == "nan" else float.fromhex(rec["omega"])
        if rec["code"] == "ValueError":
            with pytest.raises(S.SatsError, match="bad omega"):
                S.tabcode_from_angle(om)
        else:
            assert S.tabcode_from_angle(om) == rec["code"], om
        seen.add(rec["code"])
    assert seen == {"PE", "PD", "RD", "RT", "OT", "OS", "LS", "LE", "ValueError"}
    assert S.tabcode_from_angle(math.pi) == "OT" and S.tabcode_from_angle(0.0) == "PE" and S.tabcode_from_angle(-0.0) == "PE"


def test_relative_angle_matches_the_reference(f3):
    """Floating point: the reference goes through numpy.linalg.det for its cross products, this library through the
    plain 2 x 2 formula, so agreement is to 1e-9 rad (observed < 1e-12), and None cases must be None."""
    worst = 0.0
    for rec in f3["pairs"]:
        got = S.relative_angle(rec["c1"], rec["d1"], rec["c2"], rec["d2"])
        if rec["omega"] is None:
            assert got is None
            continue
        want = float.fromhex(rec["omega"])
        assert got is not None and abs(got - want) < 1e-9, (rec, got, want)
        worst = max(worst, abs(got - want))
    assert worst < 1e-9


def test_build_structure_writes_the_reference_text(f3, tmp_path):
    """A 14-SSE structure from random axes (one parallel pair -> '??', one distance > 99.9 A -> clamped): the database text
    of the built structure must be byte-identical to the text assembled from the reference's functions."""
    st = f3["structure"]
    q = S.build_structure(st["name"], st["types"], st["centroid"], st["dircos"])
    assert len(q) == 1 and q.order(0) == st["n"] and q.name(0) == st["name"]
    q.write_ascii(tmp_path / "q.ascii")
    assert (tmp_path / "q.ascii").read_text() == st["ascii"]
    tab, dm = q.get(0)
    assert (tab == tab.T).all() and (dm == dm.T).all() and tab[9, 2] == 0x44 and dm[13, 0] == np.float32(99.9)
    # and it parses back to itself through the database reader
    back = S.Database.parse_ascii(st["ascii"])
    t2, d2 = back.get(0)
    assert np.array_equal(tab, t2) and np.array_equal(dm, d2)


def test_build_structure_rejects_bad_input():
    c = np.zeros((2, 3)); d = np.array([[1.0, 0, 0], [0, 1.0, 0]])
    with pytest.raises(S.SatsError, match="type"):
        S.build_structure("x", [0, 7], c, d)
    with pytest.raises(S.SatsError, match="order"):
        S.build_structure("x", [0] * 112, np.zeros((112, 3)), np.ones((112, 3)))
    with pytest.raises(S.SatsError):
        S.build_structure("x", [0, 1], np.zeros((3, 3)), d)


@pytest.mark.gpu
def test_built_structure_finds_itself(f3):
    """Round trip on the GPU: a structure built from axes, searched against a database that contains it, must rank itself
    first with (nearly) the maximum 2 * C(n, 2) any matching can reach (zeta('??', '??') is +2 as well: both letters
    agree), through a map that matches SSEs to themselves; the reported score is the full score of the reported map."""
    st = f3["structure"]
    q = S.build_structure(st["name"], st["types"], st["centroid"], st["dircos"])
    base = S.Database.read_packed(GOLDEN / "small586.satsdb")
    tabs, dms, names = [], [], []
    for i in range(0, 120):
        t, d = base.get(i)
        tabs.append(t); dms.append(d); names.append(base.name(i))
    t, d = q.get(0)
    tabs.insert(37, t); dms.insert(37, d); names.insert(37, "f3gold")
    db = S.Database.from_structures(names, tabs, dms)
    sr = S.Searcher(db, 0)
    sc, mp = sr.search(q, S.default_params(lorder=1, lsoln=1, restarts=512, seed=3))
    n = st["n"]
    assert sc[0].argmax() == 37 and n * (n - 1) * 0.8 <= sc[0, 37] <= n * (n - 1)
    m = mp[0, 37, :n]
    assert all(j == i for i, j in enumerate(m) if j >= 0)
    k = int((m >= 0).sum())
    assert sc[0, 37] == k * (k - 1)                         # every matched pair scores +2
    sr.close()


def test_fit_axis_matches_the_reference(f3):
    """Axes from C-alpha traces (helices and strands of 1..20 residues, random pose, noisy): the reference takes the first right
    singular vector of the centred midpoints (LAPACK SVD), this library the dominant eigenvector of their scatter matrix
    (Jacobi) -- agreement to 1e-9 in every component of the direction cosines and the centroid, same orientation (N -> C),
    and None exactly where the reference gives None (helix < 3 residues, strand < 2)."""
    worst = 0.0
    none = 0
    for rec in f3["axes"]:
        got = S.fit_axis(rec["type"], rec["ca"])
        if rec["dircos"] is None:
            assert got is None, rec
            none += 1
            continue
        assert got is not None
        d, c = got
        worst = max(worst, float(np.abs(d - np.array(rec["dircos"])).max()), float(np.abs(c - np.array(rec["centroid"])).max()))
        assert abs(np.linalg.norm(d) - 1.0) < 1e-12
    assert worst < 1e-9 and none == 10, (worst, none)


def test_structure_from_ca_traces_writes_the_reference_text(f3, tmp_path):
    """Nine SSEs given as C-alpha traces, one of them a two-residue helix without an axis ('??' codes and 0.000 distances
    against everything, as the reference's None propagates): the database text must be byte-identical to the one assembled from
    the reference's fit_axis / relative_angle / angle_to_tabcode."""
    rec = f3["from_ca"]
    q = S.build_structure_from_ca(rec["name"], rec["types"], rec["traces"])
    q.write_ascii(tmp_path / "q.ascii")
    assert (tmp_path / "q.ascii").read_text() == rec["ascii"]
    tab, dm = q.get(0)
    assert (tab[4, [0, 1, 2, 3, 5, 6, 7, 8]] == 0x44).all() and (dm[4, [0, 1, 2, 3, 5, 6, 7, 8]] == 0).all()
    with pytest.raises(S.SatsError):
        S.build_structure_from_ca("x", [0, 1], [np.zeros((3, 3))])
