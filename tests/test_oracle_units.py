"""Unit checks of the oracle's building blocks: zeta, the incremental-delta invariant the reference
asserts under TESTING (cudaSaTabsearch_kernel.cu:1105-1134), Philox known answers, XORWOW stepping."""
import numpy as np

from _refio import Structure


def test_zeta_table(oracle):
    z = oracle.lib.sats_oracle_zeta
    for hx in range(5):
        for lx in range(5):
            for hy in range(5):
                for ly in range(5):
                    want = 2 if (hx == hy and lx == ly) else 1 if (hx == hy or lx == ly) else -2
                    assert z((hx << 4) | lx, (hy << 4) | ly) == want


def test_delta_equals_full_rescore(oracle, fixtures):
    rng = np.random.default_rng(7)
    q = fixtures["queries_by_name"]["D2PHLB1"]
    db = [s for s in fixtures["small586"] if s.n >= 12][:40]
    for e in db:
        for _ in range(25):
            m = np.full(q.n, -1, np.int32)
            js = rng.permutation(e.n)[: rng.integers(0, min(q.n, e.n) + 1)]
            ks = rng.permutation(q.n)[: len(js)]
            m[ks] = js
            i = int(rng.integers(0, q.n))
            free = [j for j in range(e.n) if j not in set(m.tolist())]
            to = int(rng.choice(free)) if free and rng.random() < 0.8 else -1
            frm = int(m[i])
            base = oracle.full_score(q, e, m)
            d = oracle.delta_score(q, e, m, i, frm, to)
            m2 = m.copy(); m2[i] = to
            assert base + d == oracle.full_score(q, e, m2)


def test_philox_known_answers(oracle):
    # Random123 kat_vectors, philox4x32-10
    assert oracle.philox([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert oracle.philox([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert oracle.philox([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_xorwow_stream0_is_seed_scramble(oracle):
    """Stream 0 is the salted-seed state itself (no jump); seed 0 stepping follows Marsaglia's recurrence."""
    st = oracle.xorwow_states(4, seed=1234)
    s0 = (1234 ^ 0xaad26b49) & 0xffffffff
    t0 = (1099087573 * s0) & 0xffffffff
    t1 = (2591861531 * 0xf7dcefdd) & 0xffffffff
    want = [(6615241 + t1 + t0) & 0xffffffff, (123456789 + t0) & 0xffffffff, 362436069 ^ t0,
            (521288629 + t1) & 0xffffffff, 88675123 ^ t1, (5783321 + t0) & 0xffffffff]
    assert st[0].tolist() == want
    assert len({tuple(r) for r in st.tolist()}) == 4          # distinct subsequences
    assert all(r[0] == want[0] for r in st.tolist())          # the Weyl word does not move under 2^67 jumps
    # python re-statement of one step
    s = st[1].copy()
    x = [int(v) for v in s]
    t = x[1] ^ (x[1] >> 2)
    x[1:5] = x[2:6]
    x[5] = (x[5] ^ ((x[5] << 4) & 0xffffffff)) ^ (t ^ ((t << 1) & 0xffffffff))
    x[0] = (x[0] + 362437) & 0xffffffff
    got = oracle.lib.sats_oracle_xorwow_next(s.ctypes.data)
    assert got == (x[5] + x[0]) & 0xffffffff and s.tolist() == x


def test_stream_layout_v2_candidate_draw_is_uniform_given_the_pick(oracle):
    """Production streams, layout v2: one 32-bit word picks the query SSE (its leading bits decide) and, Fibonacci-hashed
    (word * 0x9E3779B9), also draws the candidate.  For that to be sound the hashed draw must be uniform CONDITIONAL on the
    pick -- checked here for every pick value of a small, the bench and the largest query: chi-square of the candidate index
    over 2..7 candidates, and the correlation of the two uniforms."""
    pick = oracle.lib.sats_oracle_pick_from_bits
    rng = np.random.default_rng(2024)
    x = rng.integers(0, 2**32, 400000, dtype=np.uint64)
    u1 = (x.astype(np.float32) * np.float32(2.3283064e-10) + np.float32(1.1641532e-10)).astype(np.float64)
    h = (x * np.uint64(0x9E3779B9)) & np.uint64(0xffffffff)
    u2 = (h.astype(np.float32) * np.float32(2.3283064e-10) + np.float32(1.1641532e-10)).astype(np.float64)
    assert abs(np.corrcoef(u1, u2)[0, 1]) < 0.01
    for n1 in (8, 19, 111):
        i = np.minimum(((u1 - 1.1e-7) * n1).astype(np.int64), n1 - 1)
        i = np.maximum(i, 0)
        for k in (0, n1 // 2, n1 - 1):
            assert pick(int(x[k]), n1) == int(i[k])                      # the vectorised pick is the oracle's
        for ncand in (2, 3, 7):
            c = np.maximum(((u2 - 1.1e-7) * ncand).astype(np.int64), 0)
            for iv in range(n1):
                sel = c[i == iv]
                if len(sel) < 1500:
                    continue
                obs = np.bincount(sel, minlength=ncand).astype(np.float64)
                exp = len(sel) / ncand
                chi2 = ((obs - exp) ** 2 / exp).sum()
                assert chi2 < 40.0, (n1, ncand, iv, obs)                # 6 degrees of freedom at most: p < 1e-6 if it failed
    # and the seeding bits: one bit per query SSE, all 128 of a block usable
    ctr, key = [0x80000000, 5, 77, 3], [1234, 0]
    words = oracle.philox(ctr, key)
    assert len(words) == 4 and all(0 <= w < 2**32 for w in words)
