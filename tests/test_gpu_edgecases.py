"""Randomised edge-case parity: synthetic structures that stress the corners the real fixtures do not reach --
orders 1..111 (mask-word boundaries 32/33, 64/65, 96/97), all four SSE types, the full 5x5 code alphabet incl. '?',
distance differences of exactly 4.000 (the gate's boundary), NaN distances, single-restart and odd restart counts."""
import numpy as np
import pytest

import cuda_satabsearch_b200 as S
from _refio import Structure

pytestmark = pytest.mark.gpu


def random_structure(rng, name, n, grid=0.5):
    tab = np.zeros((n, n), np.uint8)
    dm = np.zeros((n, n), np.float32)
    types = rng.choice(4, n, p=[0.5, 0.35, 0.05, 0.10]).astype(np.uint8)
    for i in range(n):
        tab[i, i] = types[i]
        dm[i, i] = types[i]
        for j in range(i):
            code = (int(rng.integers(0, 5)) << 4) | int(rng.integers(0, 5))
            tab[i, j] = tab[j, i] = code
            d = np.float32(rng.integers(0, 60) * grid)        # coarse grid => many differences of exactly 4.0
            if rng.random() < 0.01:
                d = np.float32(np.nan)
            dm[i, j] = dm[j, i] = d
    return Structure(name, tab, dm)


def to_db(structs):
    return S.Database.from_structures([s.name for s in structs], [s.tab for s in structs], [s.dmat for s in structs])


@pytest.fixture(scope="module")
def synth():
    rng = np.random.default_rng(12345)
    orders = [1, 1, 2, 3, 5, 8, 13, 17, 31, 32, 33, 40, 63, 64, 65, 80, 95, 96, 97, 104, 110, 111] + \
             [int(x) for x in rng.integers(2, 60, 40)]
    ents = [random_structure(rng, "e%05d" % k, n) for k, n in enumerate(orders)]
    queries = [random_structure(rng, "q%05d" % k, n) for k, n in enumerate([1, 2, 7, 19, 32, 33, 65, 111])]
    return ents, queries


@pytest.mark.parametrize("lorder,lsoln,restarts", [(True, True, 128), (False, True, 64), (True, False, 1), (False, False, 33),
                                                   (True, True, 257)])
def test_philox_mode_on_synthetic_corners(synth, oracle, lorder, lsoln, restarts):
    ents, queries = synth
    sr = S.Searcher(to_db(ents), 0)
    p = S.default_params(lorder=lorder, lsoln=lsoln, restarts=restarts, seed=31337)
    got_s, got_m = sr.search(to_db(queries), p, query_index_base=5)
    for k, q in enumerate(queries):
        want_s, want_m = oracle.search_philox(q, ents, lorder=lorder, lsoln=lsoln, restarts=restarts, seed=31337,
                                              query_index=5 + k)
        assert np.array_equal(got_s[k], want_s), (k, q.n, np.nonzero(got_s[k] != want_s)[0][:8])
        if lsoln:
            assert np.array_equal(got_m[k, :, :q.n], want_m[:, :q.n]), (k, q.n)
    sr.close()


@pytest.mark.parametrize("lorder,lsoln,restarts", [(True, True, 128), (False, False, 200)])
def test_xorwow_mode_on_synthetic_corners(synth, oracle, lorder, lsoln, restarts):
    ents, queries = synth
    qs = [queries[2], queries[5], queries[7]]
    states = oracle.xorwow_states()
    sr = S.Searcher(to_db(ents), 0)
    got = np.full((3, len(ents)), -1, np.int32)
    gotm = np.full((3, len(ents), S.MAP_STRIDE), -1, np.int32)
    for pool in (S.POOL_SMALL, S.POOL_LARGE):
        p = S.default_params(lorder=lorder, lsoln=lsoln, restarts=restarts, rng_mode=S.RNG_XORWOW_GRID, pool=pool)
        sr.search(to_db(qs), p, scores=got, maps=gotm if lsoln else None)
    small = [k for k, s in enumerate(ents) if s.n <= 96]
    large = [k for k, s in enumerate(ents) if s.n > 96]
    for ids in (small, large):
        for k, q in enumerate(qs):
            ws, wm = oracle.search_xorwow_grid(q, [ents[i] for i in ids], states, lorder=lorder, lsoln=lsoln, restarts=restarts)
            assert np.array_equal(got[k, ids], ws), (k, q.n)
            if lsoln:
                assert np.array_equal(gotm[k, ids, :q.n], wm[:, :q.n])
    assert np.array_equal(sr.xorwow_states(), states)
    sr.close()


def test_empty_database_and_empty_pool(synth):
    ents, queries = synth
    empty = S.Database.parse_ascii("")
    sr = S.Searcher(empty, 0)
    sc, _ = sr.search(to_db(queries[:2]), S.default_params(restarts=32))
    assert sc.shape == (2, 0)
    sr.close()
    small_only = [s for s in ents if s.n <= 96][:10]
    sr = S.Searcher(to_db(small_only), 0)
    sc = np.full((1, 10), -7, np.int32)
    sr.search(to_db(queries[:1]), S.default_params(restarts=32, pool=S.POOL_LARGE), scores=sc)
    assert (sc == -7).all()                                   # nothing in the large pool: outputs untouched
    sr.close()


@pytest.mark.parametrize("lorder,lsoln,restarts", [(True, True, 128), (False, True, 40), (True, False, 96)])
def test_database_orders_above_111(synth, oracle, lorder, lsoln, restarts, tmp_path):
    """database structures of order 112..128 (the reference drops them, parsetableaux.c:457-465) are
    searched when the caller opts in; results still equal the oracle's, whose working arrays hold 128 SSEs."""
    _, queries = synth
    rng = np.random.default_rng(777)
    ents = [random_structure(rng, "L%05d" % k, n) for k, n in enumerate([112, 113, 120, 127, 128, 128, 30, 96, 111, 5])]
    qs = [queries[3], queries[5], queries[7]]                 # n1 = 19, 33, 111
    sr = S.Searcher(to_db(ents), 0)
    p = S.default_params(lorder=lorder, lsoln=lsoln, restarts=restarts, seed=99)
    got_s, got_m = sr.search(to_db(qs), p)
    for k, q in enumerate(qs):
        want_s, want_m = oracle.search_philox(q, ents, lorder=lorder, lsoln=lsoln, restarts=restarts, seed=99, query_index=k)
        assert np.array_equal(got_s[k], want_s), (k, q.n, got_s[k], want_s)
        if lsoln:
            assert np.array_equal(got_m[k, :, :q.n], want_m[:, :q.n]), (k, q.n)
    sr.close()
    # a query of more than 111 SSEs is refused (the SSE-map row stride is the reference's MAXDIM)
    sr = S.Searcher(to_db(ents[:2]), 0)
    with pytest.raises(S.SatsError, match="limited to 111"):
        sr.search(to_db([ents[0]]), S.default_params(restarts=32))
    sr.close()
    # the ASCII route: default reader drops them, the opt-in reader keeps them
    from _refio import write_ascii_db
    small_d = [Structure(s.name, s.tab, np.nan_to_num(np.minimum(s.dmat, np.float32(99.999)))) for s in ents]
    write_ascii_db(tmp_path / "big.ascii", small_d)
    assert len(S.Database.read_ascii(tmp_path / "big.ascii")) == 4
    assert len(S.Database.read_ascii(tmp_path / "big.ascii", max_order=S.MAXDIM_EXT)) == 10


def test_device_results_and_overlapped_collect(fixtures):
    """The multi-GPU host calls of round 2: sats_device_init, sats_search_collect_begin (copy-back enqueued, no wait), and the
    device-resident results a caller gathers with its own collective (sats_search_device_results + sats_searcher_entry_index):
    read straight from the device pointer they must be the scores collect() delivers, column k = original entry index[k]."""
    import torch
    ents = fixtures["small586"][:200]
    qs = [fixtures["queries_by_name"][n] for n in ("D2PHLB1", "D1UBIA_", "SHEETBC")]      # two size classes -> slots are reordered
    db = S.Database.from_structures([s.name for s in ents], [s.tab for s in ents], [s.dmat for s in ents])
    q = S.Database.from_structures([s.name for s in qs], [s.tab for s in qs], [s.dmat for s in qs])
    assert S.lib().sats_device_init(0) == 0 and S.lib().sats_device_init(99) < 0
    shards = [S.Searcher(db, 0, r, 2) for r in range(2)]
    p = S.default_params(lorder=1, lsoln=0, restarts=64, seed=8)
    want, _ = S.Searcher(db, 0).search(q, p)
    merged = np.full_like(want, -1)
    via_device = np.full_like(want, -1)
    for sr in shards:
        sr.upload(q)
        sr.launch(p, 0)
    for sr in shards:
        sr.collect_begin()
    for sr in shards:
        sr.collect(scores=merged)
        sr.sync()
        ptr, nq, ne, slots = sr.device_results()
        assert nq == 3 and ne == sr.entries and sorted(slots.tolist()) == [0, 1, 2]

        class Dev:
            __cuda_array_interface__ = {"shape": (nq * ne,), "typestr": "<i4", "data": (ptr, False), "version": 2}
        rows = torch.as_tensor(Dev(), device="cuda:0").cpu().numpy().reshape(nq, ne)
        idx = sr.entry_index()
        orders = np.array([ents[i].n for i in idx])
        assert (np.diff(orders) <= 0).all()                      # device order = decreasing structure order
        for slot in range(nq):
            via_device[slots[slot], idx] = rows[slot]
    assert np.array_equal(merged, want) and np.array_equal(via_device, want)
    for sr in shards:
        sr.close()
