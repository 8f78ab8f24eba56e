"""BASELINE.json configurations at (or near) their full sizes: oracle comparison where the oracle finishes in seconds,
size-independent properties at full size, and the statistical agreement of production mode with the reference."""
import numpy as np
import pytest

import cuda_satabsearch_b200 as S
from _refio import GOLDEN, Structure

pytestmark = pytest.mark.gpu


def structures_of(db, idx):
    out = []
    for i in idx:
        t, d = db.get(int(i))
        out.append(Structure(db.name(int(i)), t, d))
    return out


def as_db(structs):
    return S.Database.from_structures([s.name for s in structs], [s.tab for s in structs], [s.dmat for s in structs])


@pytest.fixture(scope="module")
def base():
    return S.Database.read_packed(GOLDEN / "small586.satsdb")


def test_config2_astral_scale_validation_mode(base, fixtures, oracle):
    """configs[1]: d1ubia_ vs synthetic ASTRAL-scale db (14 297 structures, size-sorted), validation-RNG mode:
    (score, map) of every entry must equal the oracle's reference-grid run."""
    db = base.bootstrap(14297, 20240501, True)
    q = fixtures["queries_by_name"]["D1UBIA_"]
    sr = S.Searcher(db, 0)
    p = S.default_params(lorder=1, lsoln=1, restarts=128, rng_mode=S.RNG_XORWOW_GRID, seed=1234)
    sc, mp = sr.search(as_db([q]), p)
    ents = structures_of(db, range(len(db)))
    ws, wm = oracle.search_xorwow_grid(q, ents, oracle.xorwow_states(), lorder=True, lsoln=True, restarts=128)
    assert np.array_equal(sc[0], ws)
    assert np.array_equal(mp[0, :, :q.n], wm[:, :q.n])
    sr.close()


def test_config3_query_list_200_queries(base, fixtures, oracle):
    """configs[2]: 200 queries drawn from the synthetic 15k db (seed 200) searched in ONE batched call; a random
    sample of (query, entry) pairs is checked against the oracle, every score against its upper bound."""
    db = base.bootstrap(14297, 20240501, True)
    rng = np.random.default_rng(200)
    qidx = rng.choice(len(db), 200, replace=False).astype(np.int32)
    queries = db.select(qidx)
    sr = S.Searcher(db, 0)
    p = S.default_params(lorder=1, lsoln=0, restarts=128, seed=4242)
    sc, _ = sr.search(queries, p, query_index_base=0)
    assert sc.shape == (200, len(db)) and sc.min() > np.iinfo(np.int32).min
    orders = db.orders().astype(np.int64)
    for k in range(200):
        n1 = int(orders[qidx[k]])
        mn = np.minimum(n1, orders)
        assert np.all(sc[k] <= mn * (mn - 1))                      # <= 2 * C(min(n1, n2), 2)
    for k in (0, 57, 199):
        pick = np.sort(rng.choice(len(db), 250, replace=False)).astype(np.int32)
        q = structures_of(db, [qidx[k]])[0]
        ws, _ = oracle.search_philox(q, structures_of(db, pick), entry_ids=pick, lorder=True, lsoln=False, restarts=128,
                                     seed=4242, query_index=k)
        assert np.array_equal(sc[k, pick], ws)
    sr.close()


def test_config4_sheet_query_unordered_maps_1024_restarts(base, fixtures, oracle):
    """configs[3]: SHEETBC (n1 = 9), LORDER=F LSOLN=T, 1024 restarts, synthetic db: scores and SSE maps."""
    db = base.bootstrap(3000, 20240503, True)
    q = fixtures["queries_by_name"]["SHEETBC"]
    sr = S.Searcher(db, 0)
    p = S.default_params(lorder=0, lsoln=1, restarts=1024, seed=9)
    sc, mp = sr.search(as_db([q]), p)
    pick = np.arange(0, 3000, 10, dtype=np.int32)
    ws, wm = oracle.search_philox(q, structures_of(db, pick), entry_ids=pick, lorder=False, lsoln=True, restarts=1024, seed=9)
    assert np.array_equal(sc[0, pick], ws)
    assert np.array_equal(mp[0, pick, :q.n], wm[:, :q.n])
    # every reported map must be a valid one-to-one, type-respecting matching whose full score is the reported score
    ents = structures_of(db, pick)
    for e, s in zip(pick, ents):
        m = mp[0, e, :q.n]
        used = m[m >= 0]
        assert len(set(used.tolist())) == len(used) and (used < s.n).all()
        assert all(q.tab[i, i] == s.tab[j, j] for i, j in enumerate(m) if j >= 0)
        assert oracle.full_score(q, s, m) == sc[0, e]
    sr.close()


def test_config5_full_size_properties(base, fixtures):
    """configs[4] at full size (100 000 structures, D2PHLB1, 128 restarts): deterministic, invariant under
    sharding (2 shards merged == unsharded), bounded."""
    db = base.bootstrap(100000, 20240502, True)
    q = as_db([fixtures["queries_by_name"]["D2PHLB1"]])
    p = S.default_params(lorder=1, lsoln=0, restarts=128, seed=1234)
    sr = S.Searcher(db, 0)
    a, _ = sr.search(q, p)
    b, _ = sr.search(q, p)
    assert np.array_equal(a, b)
    sr.close()
    merged = np.full_like(a, np.iinfo(np.int32).min)
    for r in range(2):
        sh = S.Searcher(db, 0, r, 2)
        sh.search(q, p, scores=merged)
        sh.close()
    assert np.array_equal(merged, a)
    orders = db.orders().astype(np.int64)
    mn = np.minimum(19, orders)
    assert np.all(a[0] <= mn * (mn - 1))


@pytest.mark.parametrize("lsoln", [False, True])
def test_config5_headline_db_against_oracle(base, fixtures, oracle, lsoln):
    """configs[4], the workload the bench line is quoted on: D2PHLB1 vs the 100 000-structure db, 128 restarts.  A
    size-stratified sample of 320 entries (every 312th of the size-sorted db, plus the largest ones) is bit-compared with
    the oracle -- scores, and SSE maps in the LSOLN = T case -- out of the full-size run."""
    db = base.bootstrap(100000, 20240502, True)
    q = fixtures["queries_by_name"]["D2PHLB1"]
    sr = S.Searcher(db, 0)
    p = S.default_params(lorder=1, lsoln=int(lsoln), restarts=128, seed=1234)
    sc, mp = sr.search(as_db([q]), p)
    sr.close()
    pick = np.unique(np.concatenate([np.arange(0, 100000, 312), np.arange(99990, 100000)])).astype(np.int32)
    assert len(pick) >= 320
    ws, wm = oracle.search_philox(q, structures_of(db, pick), entry_ids=pick, lorder=True, lsoln=lsoln, restarts=128, seed=1234)
    assert np.array_equal(sc[0, pick], ws)
    if lsoln:
        assert np.array_equal(mp[0, pick, :q.n], wm[:, :q.n])


def test_production_mode_agrees_statistically_with_reference(base, fixtures, golden, oracle):
    """North star, part 2: Philox streams differ from drand48, so per-entry scores differ run to run, but the ranking
    quality must be the same.  Truth = the reference's own converged run (its captured 2013 job, 4096 restarts).
    Eight production runs (different seeds) are compared with eight reference-path runs (oracle in drand48 mode, the
    first being the golden reference output itself): mean top-50 overlap and mean ROC AUC must agree within the
    run-to-run spread."""
    blocks = golden["captured"]["cpu_2013_d2phlb1_r4096"]["blocks"]
    ents = fixtures["small586"]
    conv = {n: s for b in blocks for n, s in zip(b["names"], b["scores"])}
    truth = np.array([conv[s.name] for s in ents])
    q = fixtures["queries_by_name"]["D2PHLB1"]

    def top(x, k=50):
        return set(np.argsort(-x, kind="stable")[:k].tolist())

    def auc(score, positives):
        pos, neg = score[positives], score[~positives]
        return ((pos[:, None] > neg[None, :]).sum() + 0.5 * (pos[:, None] == neg[None, :]).sum()) / (len(pos) * len(neg))

    positives = np.zeros(len(truth), bool)
    positives[list(top(truth, 25))] = True
    seeds = [1234, 1, 2, 3, 4, 5, 6, 7]
    ref_runs = []
    for sd in seeds:
        oracle.srand48(sd)
        ref_runs.append(oracle.search_drand48(q, ents, True, False, 128)[0])
    assert ref_runs[0].tolist() == golden["cases"]["d2phlb1_small_r128"]["blocks"][0]["scores"]
    sr = S.Searcher(base, 0)
    our_runs = [sr.search(as_db([q]), S.default_params(restarts=128, seed=sd))[0][0] for sd in seeds]
    sr.close()
    ref_auc = np.array([auc(x, positives) for x in ref_runs]); our_auc = np.array([auc(x, positives) for x in our_runs])
    ref_ov = np.array([len(top(x) & top(truth)) for x in ref_runs]); our_ov = np.array([len(top(x) & top(truth)) for x in our_runs])
    spread = max(ref_auc.std(), our_auc.std(), 0.004)
    assert abs(our_auc.mean() - ref_auc.mean()) < 3 * spread / np.sqrt(len(seeds)) + 0.003, (our_auc, ref_auc)
    assert abs(our_ov.mean() - ref_ov.mean()) < 3.0 and our_ov.min() >= 30, (our_ov, ref_ov)
    ref_mean = np.mean([x.mean() for x in ref_runs]); our_mean = np.mean([x.mean() for x in our_runs])
    assert abs(our_mean - ref_mean) < 0.1, (our_mean, ref_mean)
    assert np.corrcoef(np.mean(our_runs, 0), np.mean(ref_runs, 0))[0, 1] > 0.985


def test_device_topk_matches_host_selection(base, fixtures):
    """device-side top-k (counting select) == host selection with the same rule: score descending,
    ties by decreasing structure order then file order.  Includes a query that finds itself (a score thousands above
    the rest: the wide-range fallback) and k larger than the database."""
    db = base.bootstrap(20000, 7, True)
    qidx = np.array([3, 19999, 12000], np.int32)             # tiny, largest, mid-size structures of the db itself
    queries = db.select(qidx)
    big = fixtures["queries_by_name"]["d1twfa_"]
    qs = as_db(structures_of(queries, range(3)) + [big, fixtures["queries_by_name"]["D2PHLB1"]])
    sr = S.Searcher(db, 0)
    p = S.default_params(lorder=1, lsoln=0, restarts=128, seed=11)
    sr.upload(qs)
    sr.launch(p)
    full, _ = sr.collect()
    orders = db.orders()
    devpos = np.empty(len(db), np.int64)
    devpos[np.argsort(-orders, kind="stable")] = np.arange(len(db))          # device order of every original index
    for k in (1, 10, 500):
        idx, sc = sr.topk(k)
        for q in range(len(qs)):
            want = np.lexsort((devpos, -full[q].astype(np.int64)))[:k]
            assert idx[q].tolist() == want.tolist(), (k, q)
            assert sc[q].tolist() == full[q][want].tolist()
    small = S.Searcher(db.select(np.arange(50, dtype=np.int32)), 0)
    small.upload(qs, 4, 1)
    small.launch(p)
    idx, sc = small.topk(64)
    assert (idx[0, :50] >= 0).all() and (idx[0, 50:] == -1).all() and sorted(idx[0, :50].tolist()) == list(range(50))
    sr.close(); small.close()


def test_device_significance_cut_matches_printed_z_scores(base, fixtures):
    """sats_search_hits keeps exactly the entries whose z-score -- computed the way the result printer does
    (norm2 truncated to int at the z_gumbel call, _refio.stats) -- reaches the cut; order = device order; counts beyond the
    capacity are still reported."""
    from _refio import stats
    db = base.bootstrap(6000, 21, True)
    qs = as_db([fixtures["queries_by_name"][n] for n in ("D1UBIA_", "D2PHLB1", "d1twfa_")] + structures_of(db, [17, 5999]))
    sr = S.Searcher(db, 0)
    sr.upload(qs)
    sr.launch(S.default_params(lorder=1, lsoln=0, restarts=64, seed=5))
    full, _ = sr.collect()
    orders = db.orders()
    qn = qs.orders()
    devorder = np.argsort(-orders, kind="stable")
    for z_min, cap in ((-0.4, 6000), (1.0, 6000), (2.5, 6000), (100.0, 8), (-1e9, 100)):
        cnt, idx, sc = sr.hits(z_min, cap)
        for q in range(len(qs)):
            z = np.array([stats(int(full[q, e]), int(qn[q]), int(orders[e]))[1] for e in devorder])
            want = devorder[z >= z_min]
            assert cnt[q] == len(want), (z_min, q, cnt[q], len(want))
            got = idx[q][idx[q] >= 0]
            assert got.tolist() == want[:cap].tolist(), (z_min, q)
            assert sc[q, :len(got)].tolist() == full[q][got].tolist()
            assert (idx[q, len(got):] == -1).all()
    assert cnt.min() == 6000                                   # the last cut (z >= -1e9) keeps everything
    sr.close()
    # capacity larger than a (sharded) searcher's entry count: rows keep the caller's width
    part = S.Searcher(db, 0, 1, 40)
    part.upload(qs)
    part.launch(S.default_params(lorder=1, lsoln=0, restarts=64, seed=5))
    cnt, idx, sc = part.hits(-1e9, 1000)
    assert (cnt == part.entries).all() and part.entries < 1000
    for q in range(len(qs)):
        got = idx[q][idx[q] >= 0]
        assert len(got) == part.entries and (idx[q, part.entries:] == -1).all()
        assert sc[q, :len(got)].tolist() == full[q][got].tolist()
    part.close()


def test_streamed_hits_equal_the_dense_post_pass(base, fixtures):
    """SURVEY 8(f2), streaming form: with a cut bound, the kernels append their hits to a device list while they run;
    reading that list must give exactly what sats_search_hits() selects from the dense scores afterwards -- per query,
    same device order, same capacity rule -- while copying only the hits.  Multi-query batch, whole db and a shard,
    cuts from 'nearly nothing' to 'everything'."""
    db = base.bootstrap(6000, 21, True)
    qs = as_db([fixtures["queries_by_name"][n] for n in ("D1UBIA_", "D2PHLB1", "d1twfa_")] + structures_of(db, [17, 5999]))
    p = S.default_params(lorder=1, lsoln=0, restarts=64, seed=5)
    for shard in ((0, 1), (2, 5)):
        sr = S.Searcher(db, 0, *shard)
        sr.upload(qs)
        for z_min, cap in ((2.5, 6000), (1.0, 50), (-0.4, 6000), (-1e9, 6000)):
            sr.bind_cut(z_min)
            sr.launch(p)
            cnt, idx, sc, nbytes = sr.streamed_hits(cap)
            want = sr.hits(z_min, cap)
            assert np.array_equal(cnt, want[0]) and np.array_equal(idx, want[1]) and np.array_equal(sc, want[2]), (shard, z_min)
            assert nbytes == 4 + 8 * int(cnt.sum())                     # the counter and one 8-byte record per hit
            if z_min >= 1.0:
                assert nbytes < 0.05 * 4 * len(qs) * sr.entries           # a few per cent of the dense score matrix
        sr.bind_cut(None)
        sr.launch(p)
        with pytest.raises(S.SatsError, match="bound cut"):
            sr.streamed_hits(10)
        sr.close()
