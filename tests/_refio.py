"""Test-side helpers: an independent (numpy) reader/writer for the reference's file formats and the
ctypes binding of the CPU oracle (oracle/libsats_oracle.so).

TEST INFRASTRUCTURE ONLY -- the product never imports this module.

Formats restated from the reference (paths relative to /root/reference/nvcc_src_current):
  * ASCII database / query entries: parsetableaux.c:193-294 (parse_tableau / parse_distmatrix),
    header "%8s %d" parsetableaux.c:391; writer semantics scripts/convdb2.py:182-231.
  * query input on stdin: cudaSaTabsearch.cu:667-693.
  * result rows: cudaSaTabsearch.cu:415-453.
"""
from __future__ import annotations

import ctypes as C
import math
import os
import struct
import subprocess
from dataclasses import dataclass
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent.parent
ORACLE_DIR = REPO / "oracle"
GOLDEN = Path(__file__).resolve().parent / "golden"
MAXDIM = 111

_HI = {"P": 0, "R": 1, "O": 2, "L": 3, "?": 4}
_LO = {"E": 0, "D": 1, "S": 2, "T": 3, "?": 4}
_HI_INV = "PROL?"
_LO_INV = "EDST?"
_TYPE_TXT = ["e ", "xa", "xi", "xg"]


@dataclass
class Structure:
    name: str
    tab: np.ndarray   # (n, n) uint8, symmetric; diagonal = SSE type 0..3
    dmat: np.ndarray  # (n, n) float32, symmetric; diagonal = type as float (never read by the search)

    @property
    def n(self) -> int:
        return int(self.tab.shape[0])


def _type_code(txt: str) -> int:
    if txt[0] == "e":
        return 0
    return {"a": 1, "i": 2, "g": 3}[txt[1]]


def parse_entries(lines: list[str], pos: int = 0, maxdim: int = MAXDIM) -> tuple[list[Structure], int]:
    """Parse consecutive `name order / tableau rows / distance rows` entries starting at lines[pos]."""
    out: list[Structure] = []
    nl = len(lines)
    while pos < nl:
        while pos < nl and not lines[pos].strip():
            pos += 1
        if pos >= nl:
            break
        head = lines[pos].split()
        if len(head) != 2:
            break
        name, n = head[0][:8], int(head[1])
        pos += 1
        if pos + 2 * n > nl or any(not lines[pos + r].strip() for r in range(2 * n)):
            break                      # truncated trailing entry (the reference would read stale buffers)
        keep = n <= maxdim
        tab = np.zeros((n, n), np.uint8)
        dm = np.zeros((n, n), np.float32)
        for i in range(n):
            row = lines[pos + i]
            if keep:
                for j in range(i + 1):
                    cc = row[3 * j:3 * j + 2]
                    v = _type_code(cc) if i == j else (_HI[cc[0]] << 4) | _LO[cc[1]]
                    tab[i, j] = tab[j, i] = v
        pos += n
        for i in range(n):
            row = lines[pos + i]
            if keep:
                for j in range(i + 1):
                    dm[i, j] = dm[j, i] = np.float32(float(row[7 * j:7 * j + 7].split()[0]))
        pos += n
        if keep:
            out.append(Structure(name, tab, dm))
    return out, pos


def parse_ascii_db(path: os.PathLike | str, maxdim: int = MAXDIM) -> list[Structure]:
    with open(path, "r") as fh:
        lines = fh.read().split("\n")
    return parse_entries(lines, 0, maxdim)[0]


def parse_query_input(text: str):
    """-> (dbfile, ltype, lorder, lsoln, [Structure...]) for the stdin grammar of the non -q mode."""
    lines = text.split("\n")
    dbfile = lines[0].split()[0]
    flags = lines[1].split()
    qs, _ = parse_entries(lines, 2)
    return dbfile, flags[0] == "T", flags[1] == "T", flags[2] == "T", qs


def format_entry(s: Structure) -> str:
    """ASCII text of one entry, as scripts/convdb2.py writes it (and as every shipped fixture looks)."""
    n = s.n
    rows = ["%6s %4d" % (s.name, n)]
    for i in range(n):
        cells = []
        for j in range(i + 1):
            v = int(s.tab[i, j])
            cells.append(_TYPE_TXT[v] if i == j else _HI_INV[v >> 4] + _LO_INV[v & 15])
        rows.append(" ".join(cells) + " ")
    for i in range(n):
        rows.append(" ".join("%6.3f" % float(s.dmat[i, j]) for j in range(i + 1)) + " ")
    return "\n".join(rows) + "\n"


def write_ascii_db(path, entries: list[Structure]) -> None:
    with open(path, "w") as fh:
        fh.write("\n".join(format_entry(s) for s in entries))


def write_query_input(path, dbfile: str, lorder: bool, lsoln: bool, queries: list[Structure]) -> None:
    with open(path, "w") as fh:
        fh.write("%s\nT %s %s\n" % (dbfile, "T" if lorder else "F", "T" if lsoln else "F"))
        for k, q in enumerate(queries):
            fh.write(format_entry(q))
            if k + 1 < len(queries):
                fh.write("\n")


# ----------------------------------------------------------------------------- packed binary db ("SATSDB1")
# magic[8] | u32 count | u32 0 | u64 tri_cells | i32 order[count] | char name[count][9] | pad to 8 |
# u8 tab_tri[tri_cells] | pad to 8 | f32 dmat_tri[tri_cells]; entry e holds the lower triangle incl. the
# diagonal, row-major: cell (i, j<=i) at i*(i+1)/2 + j.  Entries are in ORIGINAL FILE ORDER.
MAGIC = b"SATSDB1\0"


def _pad8(n: int) -> int:
    return (-n) % 8


def write_packed(path, entries: list[Structure]) -> None:
    orders = np.array([s.n for s in entries], np.int32)
    tri = int(sum(n * (n + 1) // 2 for n in orders))
    names = bytearray(9 * len(entries))
    for k, s in enumerate(entries):
        b = s.name.encode()[:8]
        names[9 * k:9 * k + len(b)] = b
    tabs = np.concatenate([s.tab[np.tril_indices(s.n)] for s in entries]).astype(np.uint8)
    dms = np.concatenate([s.dmat[np.tril_indices(s.n)] for s in entries]).astype(np.float32)
    with open(path, "wb") as fh:
        fh.write(MAGIC)
        fh.write(struct.pack("<IIQ", len(entries), 0, tri))
        fh.write(orders.tobytes())
        fh.write(bytes(names))
        fh.write(b"\0" * _pad8(4 * len(entries) + 9 * len(entries)))
        fh.write(tabs.tobytes())
        fh.write(b"\0" * _pad8(tri))
        fh.write(dms.tobytes())


def read_packed(path) -> list[Structure]:
    raw = Path(path).read_bytes()
    assert raw[:8] == MAGIC, "not a SATSDB1 file"
    count, _, tri = struct.unpack_from("<IIQ", raw, 8)
    pos = 24
    orders = np.frombuffer(raw, np.int32, count, pos); pos += 4 * count
    names = raw[pos:pos + 9 * count]; pos += 9 * count
    pos += _pad8(13 * count)
    tabs = np.frombuffer(raw, np.uint8, tri, pos); pos += tri + _pad8(tri)
    dms = np.frombuffer(raw, np.float32, tri, pos)
    out, o = [], 0
    for k in range(count):
        n = int(orders[k]); t = n * (n + 1) // 2
        il = np.tril_indices(n)
        tab = np.zeros((n, n), np.uint8); dm = np.zeros((n, n), np.float32)
        tab[il] = tabs[o:o + t]; tab.T[il] = tabs[o:o + t]
        dm[il] = dms[o:o + t]; dm.T[il] = dms[o:o + t]
        o += t
        out.append(Structure(names[9 * k:9 * k + 9].split(b"\0")[0].decode(), tab, dm))
    return out


# ----------------------------------------------------------------------------- oracle binding
def build_oracle() -> Path:
    so = ORACLE_DIR / "libsats_oracle.so"
    src = ORACLE_DIR / "sats_oracle.c"
    if not so.exists() or so.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["make", "-C", str(ORACLE_DIR), "libsats_oracle.so"], check=True, capture_output=True)
    return so


def _dense(entries: list[Structure]):
    orders = np.array([s.n for s in entries], np.int32)
    off = np.zeros(len(entries), np.int64)
    if len(entries) > 1:
        off[1:] = np.cumsum(orders[:-1].astype(np.int64) ** 2)
    tabs = np.concatenate([s.tab.ravel() for s in entries]).astype(np.uint8) if entries else np.zeros(0, np.uint8)
    dms = np.concatenate([s.dmat.ravel() for s in entries]).astype(np.float32) if entries else np.zeros(0, np.float32)
    return orders, off, np.ascontiguousarray(tabs), np.ascontiguousarray(dms)


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


class Oracle:
    """ctypes face of oracle/sats_oracle.c.  Scores -> int32[count]; maps -> int32[count, 111] or None."""

    def __init__(self):
        self.lib = C.CDLL(str(build_oracle()))
        L = self.lib
        L.sats_oracle_zeta.restype = C.c_int
        L.sats_oracle_accept_threshold.restype = C.c_float
        L.sats_oracle_accept_threshold.argtypes = [C.c_int, C.c_int]
        for f in ("norm2", "zscore", "pvalue"):
            getattr(L, "sats_oracle_" + f).restype = C.c_double
        L.sats_oracle_norm2.argtypes = [C.c_int, C.c_int, C.c_int]
        L.sats_oracle_zscore.argtypes = [C.c_double]
        L.sats_oracle_pvalue.argtypes = [C.c_double]
        L.sats_oracle_srand48.argtypes = [C.c_long]
        L.sats_oracle_xorwow_init.argtypes = [C.c_void_p, C.c_int, C.c_uint64]
        L.sats_oracle_xorwow_next.restype = C.c_uint32
        L.sats_oracle_xorwow_next.argtypes = [C.c_void_p]

    def srand48(self, seed: int = 1234):
        self.lib.sats_oracle_srand48(seed)

    def xorwow_states(self, n: int = 128 * 128, seed: int = 1234) -> np.ndarray:
        st = np.zeros((n, 6), np.uint32)
        self.lib.sats_oracle_xorwow_init(st.ctypes.data, n, seed)
        return st

    def _common(self, q: Structure, entries):
        qtab = np.ascontiguousarray(q.tab, np.uint8)
        qd = np.ascontiguousarray(q.dmat, np.float32)
        orders, off, tabs, dms = _dense(entries)
        scores = np.zeros(len(entries), np.int32)
        maps = np.full((len(entries), MAXDIM), -1, np.int32)
        args = [C.c_int(q.n), _p(qtab, C.c_uint8), _p(qd, C.c_float), C.c_int(len(entries)),
                _p(orders, C.c_int32), _p(off, C.c_int64), _p(tabs, C.c_uint8), _p(dms, C.c_float)]
        keep = (qtab, qd, orders, off, tabs, dms)
        return args, scores, maps, keep

    def search_drand48(self, q, entries, lorder=True, lsoln=False, restarts=128):
        args, scores, maps, keep = self._common(q, entries)
        rc = self.lib.sats_oracle_search_drand48(*args, C.c_int(lorder), C.c_int(lsoln), C.c_int(restarts),
                                                 _p(scores, C.c_int32), _p(maps, C.c_int32))
        assert rc == 0, rc
        return scores, (maps if lsoln else None)

    def search_xorwow_grid(self, q, entries, states, lorder=True, lsoln=False, restarts=128,
                           nblocks=128, nthreads=128):
        assert states.shape[0] >= nblocks * nthreads and states.dtype == np.uint32
        args, scores, maps, keep = self._common(q, entries)
        rc = self.lib.sats_oracle_search_xorwow_grid(*args, C.c_int(lorder), C.c_int(lsoln), C.c_int(restarts),
                                                     C.c_void_p(states.ctypes.data), C.c_int(nblocks),
                                                     C.c_int(nthreads), _p(scores, C.c_int32), _p(maps, C.c_int32))
        assert rc == 0, rc
        return scores, (maps if lsoln else None)

    def search_philox(self, q, entries, entry_ids=None, lorder=True, lsoln=False, restarts=128,
                      seed=1234, query_index=0):
        args, scores, maps, keep = self._common(q, entries)
        ids = None if entry_ids is None else np.ascontiguousarray(entry_ids, np.int32)
        rc = self.lib.sats_oracle_search_philox(*args, None if ids is None else _p(ids, C.c_int32),
                                                C.c_int(lorder), C.c_int(lsoln), C.c_int(restarts),
                                                C.c_uint64(seed), C.c_uint32(query_index),
                                                _p(scores, C.c_int32), _p(maps, C.c_int32))
        assert rc == 0, rc
        return scores, (maps if lsoln else None)

    def philox(self, ctr, key):
        c = (C.c_uint32 * 4)(*ctr); k = (C.c_uint32 * 2)(*key); o = (C.c_uint32 * 4)()
        self.lib.sats_oracle_philox4x32_10(c, k, o)
        return list(o)

    def full_score(self, q, e, m):
        m = np.ascontiguousarray(m, np.int32)
        return self.lib.sats_oracle_full_score(q.n, _p(np.ascontiguousarray(q.tab), C.c_uint8),
                                               _p(np.ascontiguousarray(q.dmat), C.c_float), e.n,
                                               _p(np.ascontiguousarray(e.tab), C.c_uint8),
                                               _p(np.ascontiguousarray(e.dmat), C.c_float), _p(m, C.c_int32))

    def delta_score(self, q, e, m, i, frm, to):
        m = np.ascontiguousarray(m, np.int32)
        return self.lib.sats_oracle_delta_score(q.n, _p(np.ascontiguousarray(q.tab), C.c_uint8),
                                                _p(np.ascontiguousarray(q.dmat), C.c_float), e.n,
                                                _p(np.ascontiguousarray(e.tab), C.c_uint8),
                                                _p(np.ascontiguousarray(e.dmat), C.c_float), _p(m, C.c_int32),
                                                i, frm, to)


# ----------------------------------------------------------------------------- reference-style output
GUMBEL_A = 0.3780327676087335
GUMBEL_B = 0.3582596175507505
_EULER = 0.5772156649015328606


def stats(score: int, n1: int, n2: int):
    norm2 = 2.0 * score / float(n1 + n2)
    z = (int(norm2) - (GUMBEL_A + GUMBEL_B * _EULER)) / ((math.pi / math.sqrt(6.0)) * GUMBEL_B)
    p = 1 - math.exp(-math.exp(-((math.pi / math.sqrt(6.0)) * z + _EULER)))
    return norm2, z, p


def render_pool(qname, qn, dbfile, lorder, lsoln, entries, scores, maps) -> str:
    """Text the reference prints for one (query, pool): three '#' lines then one row per entry."""
    out = ["# cudaSaTabsearch LTYPE = T LORDER = %s LSOLN = %s" % ("T" if lorder else "F", "T" if lsoln else "F"),
           "# QUERY ID = %-8s" % qname, "# DBFILE = %-80s" % dbfile]
    for e, s in enumerate(entries):
        n2s, z, p = stats(int(scores[e]), qn, s.n)
        out.append("%-8s %d %g %g %g" % (s.name, int(scores[e]), n2s, z, p))
        if lsoln:
            for k in range(qn):
                if maps[e, k] >= 0:
                    out.append("%3d %3d" % (k + 1, maps[e, k] + 1))
    return "\n".join(out) + "\n"


def split_pools(entries, threshold: int = 96):
    small = [s for s in entries if s.n <= threshold]
    large = [s for s in entries if s.n > threshold]
    return small, large
