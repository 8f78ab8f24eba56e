"""GPU parity tests: the CUDA path (through the C ABI of libsats.so) against the CPU oracle on the same inputs.

Bar: bit-exact best score per entry, and bit-exact SSE map when LSOLN=T, in both uniform-source modes
(Philox production streams; XORWOW 128x128 reference-grid validation streams).
"""
import os
import subprocess
import tempfile

import numpy as np
import pytest

import cuda_satabsearch_b200 as S
from _refio import REPO, render_pool, split_pools, write_ascii_db, write_query_input

pytestmark = pytest.mark.gpu


def to_db(structs):
    return S.Database.from_structures([s.name for s in structs], [s.tab for s in structs], [s.dmat for s in structs])


@pytest.fixture(scope="module")
def small586(fixtures):
    return fixtures["small586"]


@pytest.fixture(scope="module")
def searcher586(small586):
    return S.Searcher(to_db(small586), 0)


def test_xorwow_grid_init_matches_oracle(searcher586, oracle):
    """sats_xorwow_init_kernel uses cuRAND's own curand_init(1234, tid, 0); the oracle derives the same 16384
    states from GF(2) matrix powers of the one-step map.  Agreement validates both."""
    searcher586.reset_xorwow(1234)
    got = searcher586.xorwow_states()
    want = oracle.xorwow_states(128 * 128, 1234)
    assert np.array_equal(got, want)


PHILOX_CASES = [
    # query, lorder, lsoln, restarts, entry slice
    ("D1UBIA_", True, False, 128, slice(None)),
    ("D1UBIA_", True, True, 128, slice(0, 200)),
    ("D2PHLB1", True, False, 128, slice(None)),
    ("D2PHLB1", True, True, 64, slice(100, 260)),
    ("SHEETBC", False, True, 128, slice(0, 300)),
    ("SHEETBC", False, False, 256, slice(300, 420)),
    ("D1AE6H1", False, True, 100, slice(0, 150)),      # restarts not a multiple of the team width
    ("D1AE6H1", True, True, 32, slice(150, 400)),      # one-warp teams
    ("d1twfa_", True, True, 128, slice(0, 48)),        # n1 = 101: four-word query masks
    ("d1twfa_", False, True, 128, slice(40, 70)),
    ("2QP2d1", True, False, 300, slice(200, 330)),
]


@pytest.mark.parametrize("qname,lorder,lsoln,restarts,sl", PHILOX_CASES)
def test_philox_mode_matches_oracle(fixtures, small586, searcher586, oracle, qname, lorder, lsoln, restarts, sl):
    q = fixtures["queries_by_name"][qname]
    idx = np.arange(len(small586))[sl]
    sub = [small586[i] for i in idx]
    want_s, want_m = oracle.search_philox(q, sub, entry_ids=idx, lorder=lorder, lsoln=lsoln, restarts=restarts,
                                          seed=1234, query_index=7)
    p = S.default_params(lorder=lorder, lsoln=lsoln, restarts=restarts, rng_mode=S.RNG_PHILOX, seed=1234)
    got_s, got_m = searcher586.search(to_db([q]), p, query_index_base=7)
    assert np.array_equal(got_s[0, idx], want_s)
    if lsoln:
        assert np.array_equal(got_m[0, idx, :q.n], want_m[:, :q.n])


def test_philox_multiquery_batch_and_sharding(fixtures, small586, oracle):
    """Three queries in one call (grid.y), database split over 3 shard searchers on one GPU: every entry must
    get exactly the score of the unsharded oracle run -- chains are keyed by original entry index."""
    qs = [fixtures["queries_by_name"][n] for n in ("D1UBIA_", "D1AE6H1", "SHEETBC")]
    db = to_db(small586)
    p = S.default_params(lorder=True, lsoln=True, restarts=64, seed=99)
    scores = np.full((3, len(small586)), -1, np.int32)
    maps = np.full((3, len(small586), S.MAP_STRIDE), -9, np.int32)
    counts = []
    for r in range(3):
        sh = S.Searcher(db, 0, r, 3)
        counts.append(sh.entries)
        sh.search(to_db(qs), p, query_index_base=10, scores=scores, maps=maps)
        sh.close()
    assert sum(counts) == len(small586) and max(counts) - min(counts) <= 2
    for k, q in enumerate(qs):
        want_s, want_m = oracle.search_philox(q, small586, lorder=True, lsoln=True, restarts=64, seed=99,
                                              query_index=10 + k)
        assert np.array_equal(scores[k], want_s)
        assert np.array_equal(maps[k, :, :q.n], want_m[:, :q.n])


XORWOW_CASES = [
    (["D1UBIA_"], True, False, 128),
    (["D2PHLB1", "D1UBIA_"], True, True, 128),      # second query continues the streams of the first
    (["SHEETBC"], False, True, 256),
    (["D1AE6H1"], True, False, 100),                # rounded up to 128 chains, as the reference GPU does
]


@pytest.mark.parametrize("qnames,lorder,lsoln,restarts", XORWOW_CASES)
def test_xorwow_grid_mode_matches_oracle(fixtures, small586, oracle, qnames, lorder, lsoln, restarts):
    ents = small586[:300]
    qs = [fixtures["queries_by_name"][n] for n in qnames]
    states = oracle.xorwow_states(128 * 128, 1234)
    sr = S.Searcher(to_db(ents), 0)
    p = S.default_params(lorder=lorder, lsoln=lsoln, restarts=restarts, rng_mode=S.RNG_XORWOW_GRID, seed=1234)
    got_s, got_m = sr.search(to_db(qs), p)
    for k, q in enumerate(qs):
        want_s, want_m = oracle.search_xorwow_grid(q, ents, states, lorder=lorder, lsoln=lsoln, restarts=restarts)
        assert np.array_equal(got_s[k], want_s), (k, np.nonzero(got_s[k] != want_s)[0][:10])
        if lsoln:
            assert np.array_equal(got_m[k, :, :q.n], want_m[:, :q.n])
    assert np.array_equal(sr.xorwow_states(), states)      # every stream advanced by exactly the same draws
    sr.close()


def test_xorwow_pools_follow_reference_launch_order(fixtures, small586, oracle):
    """Reference GPU order: all queries on the small pool, then all queries on the large pool,
    one state grid throughout.  Pool threshold 32 gives this fixture a non-empty large pool."""
    ents = small586[:250]
    qs = [fixtures["queries_by_name"][n] for n in ("D1UBIA_", "D1AE6H1")]
    small, large = split_pools(ents, 32)
    assert large
    states = oracle.xorwow_states(128 * 128, 1234)
    sr = S.Searcher(to_db(ents), 0)
    want = np.zeros((2, len(ents)), np.int32)
    names = [s.name for s in ents]
    for pool in (small, large):
        ids = [names.index(s.name) for s in pool]
        for k, q in enumerate(qs):
            sc, _ = oracle.search_xorwow_grid(q, pool, states, restarts=128)
            want[k, ids] = sc
    got = np.full((2, len(ents)), -1, np.int32)
    for pool_id in (S.POOL_SMALL, S.POOL_LARGE):
        p = S.default_params(restarts=128, rng_mode=S.RNG_XORWOW_GRID, pool=pool_id, pool_threshold=32)
        sr.search(to_db(qs), p, scores=got)
    assert np.array_equal(got, want)
    sr.close()


@pytest.mark.parametrize("ngrid", [2, 3])
def test_xorwow_block_sharding_over_replicated_db(fixtures, small586, oracle, ngrid):
    """SURVEY 8(e), validation mode on several GPUs: the db is replicated, searcher g runs the reference blocks b with
    b % N == g (cudaSaTabsearch_kernel.cu:932 `dbi = blockIdx.x; dbi += gridDim.x`).  Merged scores and maps, and the final
    state grid taken block-wise from its owner, must equal the unsharded oracle run -- over two queries and both pools, so
    that the streams' carry-over (devStates shared by the launches, cudaSaTabsearch.cu:1051, :1232) is exercised too."""
    ents = small586[:260]
    qs = [fixtures["queries_by_name"][n] for n in ("D1UBIA_", "D2PHLB1")]
    small, large = split_pools(ents, 32)
    assert len(large) > 5
    states = oracle.xorwow_states(128 * 128, 1234)
    names = [s.name for s in ents]
    want_s = np.zeros((2, len(ents)), np.int32)
    want_m = np.full((2, len(ents), S.MAP_STRIDE), -1, np.int32)
    for pool in (small, large):
        ids = [names.index(s.name) for s in pool]
        for k, q in enumerate(qs):
            sc, mp = oracle.search_xorwow_grid(q, pool, states, lorder=True, lsoln=True, restarts=128)
            want_s[k, ids] = sc
            want_m[k, ids, :q.n] = mp[:, :q.n]
    db = to_db(ents)
    got_s = np.full((2, len(ents)), -7, np.int32)
    got_m = np.full((2, len(ents), S.MAP_STRIDE), -1, np.int32)
    searchers = [S.Searcher(db, 0) for _ in range(ngrid)]
    for pool_id in (S.POOL_SMALL, S.POOL_LARGE):
        for g, sr in enumerate(searchers):
            p = S.default_params(lorder=1, lsoln=1, restarts=128, rng_mode=S.RNG_XORWOW_GRID, pool=pool_id, pool_threshold=32,
                                 grid_rank=g, grid_count=ngrid)
            sr.search(to_db(qs), p, scores=got_s, maps=got_m)
    assert np.array_equal(got_s, want_s)
    for k, q in enumerate(qs):
        assert np.array_equal(got_m[k, :, :q.n], want_m[k, :, :q.n])
    merged = np.zeros_like(states)
    for g, sr in enumerate(searchers):
        st = sr.xorwow_states().reshape(128, 128, 6)
        merged.reshape(128, 128, 6)[g::ngrid] = st[g::ngrid]
        sr.close()
    assert np.array_equal(merged, states)


def test_xorwow_refuses_an_entry_sharded_searcher(small586, fixtures):
    """A searcher holding one part of the cost-weighted partition cannot reproduce the reference blocks' walk over the
    whole pool: validation mode must say so instead of returning non-reference scores."""
    sr = S.Searcher(to_db(small586[:64]), 0, 1, 2)
    q = to_db([fixtures["queries_by_name"]["D1UBIA_"]])
    with pytest.raises(S.SatsError, match="grid_rank"):
        sr.search(q, S.default_params(rng_mode=S.RNG_XORWOW_GRID))
    with pytest.raises(S.SatsError, match="XORWOW_GRID"):
        sr.search(q, S.default_params(accept_mode=S.ACCEPT_DEVICE_FAST))      # fast-math acceptance is a validation aid
    sr.search(q, S.default_params())                                            # production mode is what shards are for
    sr.close()


REF_BIN = REPO / "oracle" / "_ref" / "cudaSaTabsearch_ref"


@pytest.mark.skipif(not REF_BIN.exists(), reason="reference binary not built (oracle/build_ref.sh)")
def test_matches_reference_gpu_binary_on_this_box(fixtures, small586):
    """Same-box parity with the reference's own GPU build (sa_tabsearch_gpu<<<128,128>>>, cuRAND XORWOW,
    --use_fast_math): our XORWOW_GRID + DEVICE_FAST mode must print the same rows."""
    q = fixtures["queries_by_name"]["D1UBIA_"]
    with tempfile.TemporaryDirectory() as td:
        write_ascii_db(os.path.join(td, "db.ascii"), small586)
        write_query_input(os.path.join(td, "q.input"), "db.ascii", True, True, [q])
        with open(os.path.join(td, "q.input"), "rb") as fh:
            run = subprocess.run([str(REF_BIN), "-r", "128"], stdin=fh, cwd=td, capture_output=True, timeout=300)
        assert run.returncode == 0, run.stderr.decode()[-2000:]
        ref_rows = [ln for ln in run.stdout.decode().split("\n") if ln and not ln.startswith("#")]
    sr = S.Searcher(to_db(small586), 0)
    p = S.default_params(lorder=True, lsoln=True, restarts=128, rng_mode=S.RNG_XORWOW_GRID,
                         accept_mode=S.ACCEPT_DEVICE_FAST, seed=1234)
    sc, mp = sr.search(to_db([q]), p)
    ours = render_pool(q.name, q.n, "db.ascii", True, True, small586, sc[0], mp[0])
    our_rows = [ln for ln in ours.split("\n") if ln and not ln.startswith("#")]
    assert our_rows == ref_rows
    sr.close()
