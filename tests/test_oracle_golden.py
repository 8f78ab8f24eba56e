"""Pins the CPU oracle (oracle/sats_oracle.c) to the reference: every golden case in tests/golden/golden.json
was produced by the UNMODIFIED reference binary (`-c` path) or is the reference's own captured 2013 job
output; the oracle in drand48 mode must reproduce each stdout byte for byte (compared by md5 + scores)."""
import hashlib

import numpy as np
import pytest

from _refio import (GOLDEN, format_entry, parse_query_input, read_packed, render_pool, split_pools,
                    stats)

INPUT_QUERIES = {  # query ids each reference input file carries, in file order
    "d1ubia_.input": ["D1UBIA_"], "d2phlb1.input": ["D2PHLB1"], "d1ae6h1.input": ["D1AE6H1"],
    "multiquery.input": ["D1UBIA_", "D1AE6H1", "d1twfa_"], "1qlp_sheetbc.input": ["SHEETBC"],
    "d2phlb1.input3": ["D2PHLB1"], "d1twfa_.input": ["d1twfa_"],
}


def oracle_stdout(oracle, fixtures, case):
    """Replays cudaSaTabsearch.cu:1272-1310: srand48(1234); small pool for every query, then large pool."""
    entries = fixtures[case["db"]]
    queries = [fixtures["queries_by_name"][n] for n in INPUT_QUERIES[case["input"]]]
    small, large = split_pools(entries, case["pool_threshold"])
    oracle.srand48(1234)
    text, scores = "", []
    for pool in (small, large):
        if not pool:
            continue
        for q in queries:
            sc, mp = oracle.search_drand48(q, pool, case["lorder"], case["lsoln"], case["restarts"])
            text += render_pool(q.name, q.n, case["dbfile"], case["lorder"], case["lsoln"], pool, sc, mp)
            scores.append(sc)
    return text, scores


CASES = ["d1ubia_small_r128", "d1ubia_test1_default", "d1ae6h1_test2_r128", "d2phlb1_small_r128",
         "multiquery_small_r128", "sheetbc_d1qlpa_TTT_r1024", "sheetbc_d1qlpa_TFT_r1024",
         "sheetbc_small_TFT_r128", "d2phlb1_d2pq6a1_TTT_r128", "d2phlb1_small_r128_md32"]


@pytest.mark.parametrize("cid", CASES)
def test_oracle_reproduces_reference_stdout(oracle, fixtures, golden, cid):
    case = golden["cases"][cid]
    text, scores = oracle_stdout(oracle, fixtures, case)
    for got, blk in zip(scores, case["blocks"]):
        assert got.tolist() == blk["scores"]
    assert hashlib.md5(text.encode()).hexdigest() == case["stdout_md5"]


@pytest.mark.slow
def test_oracle_reproduces_2013_captured_job(oracle, fixtures, golden):
    """old/nvcc_src_cuda5/cpu_cudaSaTabsearch.o1462445: 586 entries x 4096 restarts, pools split at 32."""
    case = golden["captured"]["cpu_2013_d2phlb1_r4096"]
    text, scores = oracle_stdout(oracle, fixtures, case)
    for got, blk in zip(scores, case["blocks"]):
        assert got.tolist() == blk["scores"]
    assert hashlib.md5(text.encode()).hexdigest() == case["stdout_md5"]


def test_known_answers(golden):
    """The hand-made fixtures' documented answers: 54 + identity map, 92, 72 + 9-pair map."""
    b = golden["cases"]["d1ubia_test1_default"]["blocks"][0]
    assert b["names"] == ["d1ndda_"] and b["scores"] == [54]
    assert b["maps"][0] == [[k, k] for k in range(1, 9)]
    b = golden["cases"]["d1ae6h1_test2_r128"]["blocks"][0]
    assert b["names"] == ["d1kcul1"] and b["scores"] == [92]
    b = golden["cases"]["sheetbc_d1qlpa_TFT_r1024"]["blocks"][0]
    assert b["scores"] == [72]
    assert b["maps"][0] == [[1, 2], [2, 12], [3, 13], [4, 14], [5, 15], [6, 18], [7, 24], [8, 25], [9, 26]]


def test_ascii_writer_reproduces_reference_files(fixtures, golden):
    """packed fixture -> ASCII text must be byte-identical to the reference's shipped .ascii files."""
    for key, meta in golden["ascii_md5"].items():
        txt = "\n".join(format_entry(s) for s in fixtures[key])
        if len(txt) + 1 == meta["bytes"]:
            txt += "\n"                     # d2pq6a1.ascii ends with a blank line
        assert hashlib.md5(txt.encode()).hexdigest() == meta["md5"], key


def test_stats_match_oracle_c(oracle):
    for score, n1, n2 in [(54, 8, 8), (0, 8, 3), (-7, 19, 40), (92, 13, 12), (3, 8, 67), (-1, 4, 5)]:
        n2s, z, p = stats(score, n1, n2)
        L = oracle.lib
        assert L.sats_oracle_norm2(score, n1, n2) == n2s
        assert L.sats_oracle_zscore(n2s) == z
        assert L.sats_oracle_pvalue(z) == p
    assert "%g %g %g" % stats(54, 8, 8) == "6.75 11.7853 1.53059e-07"
