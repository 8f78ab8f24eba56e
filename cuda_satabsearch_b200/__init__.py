"""cuda_satabsearch_b200 -- ctypes shim over libsats.so, the B200-native SA tableau-search library.

The product is the C-ABI library (include/sats.h) and the drop-in `cudaSaTabsearch` CLI; this module is the thin
Python face the reference never had (its scripts/ tooling shells out to the binary,
scripts/qptabmatchstructs.sh:152-158).  It loads the in-tree build and FAILS LOUDLY if the library is missing:
there is no Python or CPU fallback for the search.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

_PKG = Path(__file__).resolve().parent
LIB_PATH = _PKG / "libsats.so"
CLI_PATH = _PKG / "bin" / "cudaSaTabsearch"

MAXDIM = 111
MAXDIM_EXT = 128
MAP_STRIDE = 111
RNG_PHILOX, RNG_XORWOW_GRID = 0, 1
ACCEPT_HOST_TABLE, ACCEPT_DEVICE_FAST = 0, 1
POOL_ALL, POOL_SMALL, POOL_LARGE = 0, 1, 2


class SatsError(RuntimeError):
    pass


class Params(C.Structure):
    _fields_ = [("lorder", C.c_int), ("lsoln", C.c_int), ("restarts", C.c_int), ("rng_mode", C.c_int),
                ("accept_mode", C.c_int), ("pool", C.c_int), ("pool_threshold", C.c_int),
                ("grid_rank", C.c_int), ("grid_count", C.c_int), ("reserved", C.c_int), ("seed", C.c_uint64)]


def _load() -> C.CDLL:
    if not LIB_PATH.exists():
        raise SatsError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                        "(make -C cuda_satabsearch_b200/csrc). There is no fallback path.")
    L = C.CDLL(str(LIB_PATH))
    vp, ci, cs = C.c_void_p, C.c_int, C.c_char_p
    P = C.POINTER
    sig = {
        "sats_last_error": (cs, []), "sats_version": (cs, []),
        "sats_db_read_ascii": (ci, [cs, P(vp)]), "sats_db_parse_ascii": (ci, [cs, C.c_size_t, P(vp)]),
        "sats_db_read_ascii_ext": (ci, [cs, ci, P(vp)]), "sats_db_parse_ascii_ext": (ci, [cs, C.c_size_t, ci, P(vp)]),
        "sats_input_parse": (ci, [cs, C.c_size_t, cs, C.c_size_t, P(ci), P(vp)]),
        "sats_idlist_parse": (ci, [cs, C.c_size_t, cs, ci]),
        "sats_db_from_arrays": (ci, [ci, vp, vp, vp, vp, vp, P(vp)]),
        "sats_db_free": (None, [vp]), "sats_db_count": (ci, [vp]), "sats_db_order": (ci, [vp, ci]),
        "sats_db_name": (cs, [vp, ci]), "sats_db_max_order": (ci, [vp]), "sats_db_get": (ci, [vp, ci, vp, vp]),
        "sats_db_find": (ci, [vp, cs]), "sats_db_select": (ci, [vp, vp, ci, P(vp)]),
        "sats_db_bootstrap": (ci, [vp, ci, C.c_uint64, ci, P(vp)]),
        "sats_db_write_ascii": (ci, [vp, cs]), "sats_db_write_packed": (ci, [vp, cs]),
        "sats_db_read_packed": (ci, [cs, P(vp)]),
        "sats_tabcode_from_angle": (ci, [C.c_double, cs]), "sats_relative_angle": (ci, [vp, vp, vp, vp, P(C.c_double)]),
        "sats_build_structure": (ci, [cs, ci, vp, vp, vp, P(vp)]),
        "sats_fit_axis": (ci, [ci, ci, vp, vp, vp]), "sats_build_structure_from_ca": (ci, [cs, ci, vp, vp, vp, P(vp)]),
        "sats_norm2": (C.c_double, [ci, ci, ci]), "sats_z_gumbel": (C.c_double, [ci, C.c_double, C.c_double]),
        "sats_pv_gumbel": (C.c_double, [C.c_double]),
        "sats_format_block": (C.c_size_t, [vp, C.c_size_t, cs, ci, cs, ci, ci, vp, vp, ci, vp, vp]),
        "sats_params_default": (None, [P(Params)]),
        "sats_searcher_create": (ci, [vp, ci, ci, ci, P(vp)]), "sats_searcher_free": (None, [vp]),
        "sats_searcher_entry_count": (ci, [vp]), "sats_searcher_device": (ci, [vp]),
        "sats_partition": (ci, [vp, ci, vp]),
        "sats_search": (ci, [vp, vp, ci, ci, P(Params), C.c_uint32, vp, vp]),
        "sats_search_upload": (ci, [vp, vp, ci, ci]),
        "sats_search_launch": (ci, [vp, P(Params), C.c_uint32, P(C.c_float)]),
        "sats_search_collect": (ci, [vp, vp, vp]), "sats_searcher_sync": (ci, [vp]),
        "sats_search_collect_begin": (ci, [vp]),
        "sats_search_device_results": (ci, [vp, P(vp), P(ci), P(ci), vp]), "sats_searcher_entry_index": (ci, [vp, vp]),
        "sats_searcher_launch_count": (C.c_longlong, [vp]),
        "sats_searcher_get_xorwow": (ci, [vp, vp]), "sats_searcher_reset_xorwow": (ci, [vp, C.c_uint64]),
        "sats_device_count": (ci, []), "sats_device_init": (ci, [ci]),
        "sats_pick_boundaries": (ci, [ci, vp]), "sats_accept_cutoffs": (ci, [vp, vp]), "sats_seed_cutoff": (C.c_uint32, []),
        "sats_score_threshold": (C.c_int32, [C.c_double, ci, ci]),
        "sats_search_topk": (ci, [vp, ci, vp, vp]),
        "sats_search_hits": (ci, [vp, C.c_double, ci, vp, vp, vp]),
        "sats_search_bind_cut": (ci, [vp, C.c_double]),
        "sats_search_streamed_hits": (ci, [vp, ci, vp, vp, vp, P(C.c_int64)]),
        "sats_results_parse": (ci, [cs, C.c_size_t, P(vp)]), "sats_results_free": (None, [vp]),
        "sats_results_blocks": (ci, [vp]), "sats_results_query": (cs, [vp, ci]), "sats_results_dbfile": (cs, [vp, ci]),
        "sats_results_flags": (ci, [vp, ci, P(ci)]), "sats_results_rows": (ci, [vp, ci]),
        "sats_results_row": (ci, [vp, ci, ci, cs, P(C.c_int32), P(C.c_double), P(C.c_double), P(C.c_double)]),
        "sats_results_map": (ci, [vp, ci, ci, vp, ci]),
        "sats_roc_auc": (C.c_double, [vp, vp, ci]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)          # AttributeError here = the library does not export what sats.h declares
        fn.restype, fn.argtypes = res, args
    return L


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        _lib = _load()
    return _lib


def _check(rc: int):
    if rc < 0:
        raise SatsError(f"libsats error {rc}: {lib().sats_last_error().decode(errors='replace')}")
    return rc


def default_params(**kw) -> Params:
    p = Params()
    lib().sats_params_default(C.byref(p))
    for k, v in kw.items():
        if not hasattr(p, k):
            raise TypeError(f"unknown search parameter {k}")
        setattr(p, k, int(v))
    return p


class Database:
    """A list of structures (a database or a query set) held by libsats in original file order."""

    def __init__(self, handle):
        self._h = C.c_void_p(handle)

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.sats_db_free(self._h)
            self._h = None

    # -- constructors
    @classmethod
    def read_ascii(cls, path, max_order: int = MAXDIM):
        h = C.c_void_p()
        _check(lib().sats_db_read_ascii_ext(os.fsencode(path), max_order, C.byref(h)))
        return cls(h.value)

    @classmethod
    def parse_ascii(cls, text: bytes | str, max_order: int = MAXDIM):
        b = text.encode() if isinstance(text, str) else text
        h = C.c_void_p()
        _check(lib().sats_db_parse_ascii_ext(b, len(b), max_order, C.byref(h)))
        return cls(h.value)

    @classmethod
    def read_packed(cls, path):
        h = C.c_void_p()
        _check(lib().sats_db_read_packed(os.fsencode(path), C.byref(h)))
        return cls(h.value)

    @classmethod
    def from_structures(cls, names, tabs, dmats):
        """names: list[str]; tabs/dmats: lists of (n, n) uint8 / float32 arrays."""
        count = len(names)
        order = np.array([t.shape[0] for t in tabs], np.int32)
        off = np.zeros(count, np.int64)
        if count > 1:
            off[1:] = np.cumsum(order[:-1].astype(np.int64) ** 2)
        nm = bytearray(9 * count)
        for k, s in enumerate(names):
            b = s.encode()[:8]
            nm[9 * k:9 * k + len(b)] = b
        nmb = (C.c_char * len(nm)).from_buffer(nm) if count else None
        t = np.ascontiguousarray(np.concatenate([np.asarray(x, np.uint8).ravel() for x in tabs])) if count else np.zeros(1, np.uint8)
        d = np.ascontiguousarray(np.concatenate([np.asarray(x, np.float32).ravel() for x in dmats])) if count else np.zeros(1, np.float32)
        h = C.c_void_p()
        _check(lib().sats_db_from_arrays(count, order.ctypes.data, C.addressof(nmb) if count else None,
                                         off.ctypes.data, t.ctypes.data, d.ctypes.data, C.byref(h)))
        return cls(h.value)

    # -- accessors
    def __len__(self):
        return lib().sats_db_count(self._h)

    def order(self, i: int) -> int:
        return _check(lib().sats_db_order(self._h, i))

    def orders(self) -> np.ndarray:
        return np.array([self.order(i) for i in range(len(self))], np.int32)

    def name(self, i: int) -> str:
        return lib().sats_db_name(self._h, i).decode()

    def names(self):
        return [self.name(i) for i in range(len(self))]

    def max_order(self) -> int:
        return lib().sats_db_max_order(self._h)

    def get(self, i: int):
        n = self.order(i)
        tab = np.zeros((n, n), np.uint8)
        dm = np.zeros((n, n), np.float32)
        _check(lib().sats_db_get(self._h, i, tab.ctypes.data, dm.ctypes.data))
        return tab, dm

    def find(self, name: str) -> int:
        return _check(lib().sats_db_find(self._h, name.encode()))

    def select(self, index) -> "Database":
        idx = np.ascontiguousarray(index, np.int32)
        h = C.c_void_p()
        _check(lib().sats_db_select(self._h, idx.ctypes.data, len(idx), C.byref(h)))
        return Database(h.value)

    def bootstrap(self, count: int, seed: int, sort_by_order: bool = True) -> "Database":
        h = C.c_void_p()
        _check(lib().sats_db_bootstrap(self._h, count, seed, int(sort_by_order), C.byref(h)))
        return Database(h.value)

    def partition(self, shard_count: int) -> np.ndarray:
        owner = np.zeros(len(self), np.int32)
        _check(lib().sats_partition(self._h, shard_count, owner.ctypes.data))
        return owner

    def write_ascii(self, path):
        _check(lib().sats_db_write_ascii(self._h, os.fsencode(path)))

    def write_packed(self, path):
        _check(lib().sats_db_write_packed(self._h, os.fsencode(path)))

    def format_block(self, query_id: str, query_order: int, dbfile: str, lorder: bool, lsoln: bool,
                     scores: np.ndarray, maps: np.ndarray | None = None, index=None) -> str:
        sc = np.ascontiguousarray(scores, np.int32)
        mp = None if maps is None else np.ascontiguousarray(maps, np.int32)
        idx = None if index is None else np.ascontiguousarray(index, np.int32)
        count = len(self) if idx is None else len(idx)
        args = (query_id.encode(), query_order, dbfile.encode(), int(lorder), int(lsoln), self._h,
                None if idx is None else idx.ctypes.data, count, sc.ctypes.data,
                None if mp is None else mp.ctypes.data)
        need = lib().sats_format_block(None, 0, *args)
        buf = C.create_string_buffer(need + 1)
        lib().sats_format_block(buf, need + 1, *args)
        return buf.raw[:need].decode()


def parse_input(text: bytes | str):
    """The reference's stdin grammar (non -q mode) -> (dbfile, ltype, lorder, lsoln, queries: Database)."""
    b = text.encode() if isinstance(text, str) else text
    dbfile = C.create_string_buffer(4096)
    flags = (C.c_int * 3)()
    h = C.c_void_p()
    _check(lib().sats_input_parse(b, len(b), dbfile, 4096, flags, C.byref(h)))
    return dbfile.value.decode(), bool(flags[0]), bool(flags[1]), bool(flags[2]), Database(h.value)


def parse_idlist(text: bytes | str):
    b = text.encode() if isinstance(text, str) else text
    cap = b.count(b"\n") + 2
    buf = C.create_string_buffer(9 * cap)
    n = _check(lib().sats_idlist_parse(b, len(b), buf, cap))
    return [buf.raw[9 * k:9 * k + 9].split(b"\0")[0].decode() for k in range(n)]


def tabcode_from_angle(omega: float) -> str:
    """scripts/pttableau.py angle_to_tabcode; raises SatsError where the reference raises ValueError."""
    buf = C.create_string_buffer(3)
    _check(lib().sats_tabcode_from_angle(float(omega), buf))
    return buf.value.decode()


def relative_angle(c_self, d_self, c_other, d_other):
    """scripts/ptnode.py PTNode.relative_angle on (centroid, direction cosines) axes; None where the reference gives None."""
    a = [np.ascontiguousarray(x, np.float64) for x in (c_self, d_self, c_other, d_other)]
    om = C.c_double(0.0)
    rc = _check(lib().sats_relative_angle(*(x.ctypes.data for x in a), C.byref(om)))
    return None if rc == 1 else om.value


def build_structure(name: str, sse_types, centroids, dircos) -> Database:
    """A one-structure Database (usable as a query) from fitted SSE axes: tableau codes from the pairwise interaxial angles,
    midpoint distances rounded as the database writer does."""
    t = np.ascontiguousarray(sse_types, np.uint8)
    c = np.ascontiguousarray(centroids, np.float64)
    d = np.ascontiguousarray(dircos, np.float64)
    if c.shape != (len(t), 3) or d.shape != (len(t), 3):
        raise SatsError("centroids and dircos must be (n, 3) arrays matching sse_types")
    h = C.c_void_p()
    _check(lib().sats_build_structure(name.encode(), len(t), t.ctypes.data, c.ctypes.data, d.ctypes.data, C.byref(h)))
    return Database(h.value)


def fit_axis(sse_type: int, ca_xyz):
    """scripts/ptnode.py fit_axis on a C-alpha trace (n_res x 3) -> (dircos, centroid) or None where the reference gives None."""
    ca = np.ascontiguousarray(ca_xyz, np.float64).reshape(-1, 3)
    d = np.zeros(3); c = np.zeros(3)
    rc = _check(lib().sats_fit_axis(int(sse_type), len(ca), ca.ctypes.data, d.ctypes.data, c.ctypes.data))
    return None if rc == 1 else (d, c)


def build_structure_from_ca(name: str, sse_types, ca_traces) -> Database:
    """A one-structure Database from the SSEs' C-alpha traces (a list of (n_res, 3) arrays, N- to C-terminus)."""
    t = np.ascontiguousarray(sse_types, np.uint8)
    traces = [np.ascontiguousarray(x, np.float64).reshape(-1, 3) for x in ca_traces]
    if len(traces) != len(t):
        raise SatsError("one C-alpha trace per SSE")
    nres = np.array([len(x) for x in traces], np.int32)
    ca = np.ascontiguousarray(np.concatenate(traces)) if traces else np.zeros((1, 3))
    h = C.c_void_p()
    _check(lib().sats_build_structure_from_ca(name.encode(), len(t), t.ctypes.data, nres.ctypes.data, ca.ctypes.data, C.byref(h)))
    return Database(h.value)


def norm2(score, n1, n2):
    return lib().sats_norm2(score, n1, n2)


def z_gumbel(x, a=None, b=None):
    L = lib()
    a = C.c_double.in_dll(L, "sats_gumbel_a").value if a is None else a
    b = C.c_double.in_dll(L, "sats_gumbel_b").value if b is None else b
    return L.sats_z_gumbel(int(x), a, b)


def pv_gumbel(z):
    return lib().sats_pv_gumbel(z)


def device_count() -> int:
    return lib().sats_device_count()


def _out_array(a: np.ndarray, shape, what: str) -> np.ndarray:
    """A caller-supplied output buffer goes to the C ABI as a bare pointer: it must be exactly what the library writes."""
    if not isinstance(a, np.ndarray) or a.dtype != np.int32 or not a.flags.c_contiguous or not a.flags.writeable \
            or tuple(a.shape) != tuple(shape):
        raise SatsError(f"{what} must be a writable C-contiguous int32 array of shape {tuple(shape)}, got "
                        f"{getattr(a, 'dtype', type(a))} {getattr(a, 'shape', '')}")
    return a


class Searcher:
    """One GPU's resident copy of (a shard of) a database plus the search entry points."""

    def __init__(self, db: Database, device: int = 0, shard_rank: int = 0, shard_count: int = 1):
        h = C.c_void_p()
        _check(lib().sats_searcher_create(db._h, device, shard_rank, shard_count, C.byref(h)))
        self._h = h
        self.db = db
        self.count = len(db)

    def close(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.sats_searcher_free(self._h)
            self._h = None

    __del__ = close

    @property
    def entries(self) -> int:
        return lib().sats_searcher_entry_count(self._h)

    @property
    def launches(self) -> int:
        return lib().sats_searcher_launch_count(self._h)

    def search(self, queries: Database, params: Params | None = None, qfirst: int = 0, qcount: int | None = None,
               query_index_base: int = 0, scores: np.ndarray | None = None, maps: np.ndarray | None = None):
        """-> (scores int32 [q, D], maps int32 [q, D, 111] or None), indexed by original db order."""
        p = params or default_params()
        qcount = len(queries) - qfirst if qcount is None else qcount
        if scores is None:
            scores = np.full((qcount, self.count), np.iinfo(np.int32).min, np.int32)
        if p.lsoln and maps is None:
            maps = np.full((qcount, self.count, MAP_STRIDE), -1, np.int32)
        _out_array(scores, (qcount, self.count), "scores")
        if p.lsoln:
            _out_array(maps, (qcount, self.count, MAP_STRIDE), "maps")
        _check(lib().sats_search(self._h, queries._h, qfirst, qcount, C.byref(p), query_index_base,
                                 scores.ctypes.data, maps.ctypes.data if p.lsoln else None))
        return scores, (maps if p.lsoln else None)

    def upload(self, queries: Database, qfirst: int = 0, qcount: int | None = None):
        qcount = len(queries) - qfirst if qcount is None else qcount
        _check(lib().sats_search_upload(self._h, queries._h, qfirst, qcount))
        self._qcount = qcount

    def launch(self, params: Params, query_index_base: int = 0, timed: bool = False):
        ms = C.c_float(0)
        _check(lib().sats_search_launch(self._h, C.byref(params), query_index_base, C.byref(ms) if timed else None))
        self._lsoln = bool(params.lsoln)
        return ms.value if timed else None

    def collect(self, scores: np.ndarray | None = None, maps: np.ndarray | None = None):
        q = self._qcount
        if scores is None:
            scores = np.full((q, self.count), np.iinfo(np.int32).min, np.int32)
        if self._lsoln and maps is None:
            maps = np.full((q, self.count, MAP_STRIDE), -1, np.int32)
        _out_array(scores, (q, self.count), "scores")
        if self._lsoln:
            _out_array(maps, (q, self.count, MAP_STRIDE), "maps")
        _check(lib().sats_search_collect(self._h, scores.ctypes.data, maps.ctypes.data if self._lsoln else None))
        return scores, (maps if self._lsoln else None)

    def sync(self):
        _check(lib().sats_searcher_sync(self._h))

    def collect_begin(self):
        _check(lib().sats_search_collect_begin(self._h))

    def device_results(self):
        """After launch(): (device pointer to int32 [qcount, entries] in device order, qcount, entries, slot -> batch position).
        For gathering shards with a collective; sync() first (the results are produced on the searcher's own stream)."""
        ptr = C.c_void_p()
        q, e = C.c_int(0), C.c_int(0)
        slots = np.zeros(max(1, getattr(self, "_qcount", 1)), np.int32)
        _check(lib().sats_search_device_results(self._h, C.byref(ptr), C.byref(q), C.byref(e), slots.ctypes.data))
        return ptr.value, q.value, e.value, slots[:q.value]

    def entry_index(self) -> np.ndarray:
        """Original db index of every resident entry, in device order (decreasing structure order)."""
        idx = np.zeros(max(1, self.entries), np.int32)
        _check(lib().sats_searcher_entry_index(self._h, idx.ctypes.data))
        return idx[:self.entries]

    def topk(self, k: int):
        """After launch(): device-side selection -> (index int32 [q, k] original db indices, scores int32 [q, k])."""
        q = self._qcount
        idx = np.full((q, k), -1, np.int32)
        sc = np.full((q, k), np.iinfo(np.int32).min, np.int32)
        _check(lib().sats_search_topk(self._h, k, idx.ctypes.data, sc.ctypes.data))
        return idx, sc

    def hits(self, z_min: float, cap: int):
        """After launch(): device-side cut at Gumbel z-score >= z_min -> (counts int32 [q], index int32 [q, cap] original
        db indices in device order, scores int32 [q, cap]); counts may exceed cap."""
        q = self._qcount
        cnt = np.zeros(q, np.int32)
        idx = np.full((q, cap), -1, np.int32)
        sc = np.full((q, cap), np.iinfo(np.int32).min, np.int32)
        _check(lib().sats_search_hits(self._h, float(z_min), cap, cnt.ctypes.data, idx.ctypes.data, sc.ctypes.data))
        return cnt, idx, sc

    def bind_cut(self, z_min: float | None):
        """Streaming hits: from now on every launch appends its hits (z-score >= z_min) to a device list; None unbinds."""
        _check(lib().sats_search_bind_cut(self._h, float("nan") if z_min is None else float(z_min)))

    def streamed_hits(self, cap: int):
        """After a launch with a bound cut: (counts, index, scores, d2h_bytes) -- the first three exactly as hits()."""
        q = self._qcount
        cnt = np.zeros(q, np.int32)
        idx = np.full((q, cap), -1, np.int32)
        sc = np.full((q, cap), np.iinfo(np.int32).min, np.int32)
        nbytes = C.c_int64(0)
        _check(lib().sats_search_streamed_hits(self._h, cap, cnt.ctypes.data, idx.ctypes.data, sc.ctypes.data, C.byref(nbytes)))
        return cnt, idx, sc, nbytes.value

    def xorwow_states(self) -> np.ndarray:
        st = np.zeros((128 * 128, 6), np.uint32)
        _check(lib().sats_searcher_get_xorwow(self._h, st.ctypes.data))
        return st

    def reset_xorwow(self, seed: int = 1234):
        _check(lib().sats_searcher_reset_xorwow(self._h, seed))


def pick_boundaries(n: int) -> np.ndarray:
    """cut[k] = smallest 32-bit draw whose SSE pick among n is >= k (validation aid, see include/sats.h)."""
    cut = np.zeros(n, np.uint32)
    _check(lib().sats_pick_boundaries(n, cut.ctypes.data))
    return cut


def accept_cutoffs():
    """-> (cut uint32 [100, 230], temps float32 [100]): the Metropolis test as integer cut-offs (validation aid)."""
    cut = np.zeros((100, 230), np.uint32)
    temps = np.zeros(100, np.float32)
    _check(lib().sats_accept_cutoffs(cut.ctypes.data, temps.ctypes.data))
    return cut, temps


def seed_cutoff() -> int:
    return int(lib().sats_seed_cutoff())


def score_threshold(z_min: float, n1: int, n2: int) -> int:
    """Smallest raw score whose printed z-score reaches z_min for sizes (n1, n2); 2**31 - 1 if none does."""
    return int(lib().sats_score_threshold(float(z_min), n1, n2))


def parse_results(text: bytes | str):
    """Five-column output -> list of blocks: dict(query, dbfile, ltype, lorder, lsoln, names, scores, norm2, z, p, maps)."""
    b = text.encode() if isinstance(text, str) else text
    h = C.c_void_p()
    _check(lib().sats_results_parse(b, len(b), C.byref(h)))
    out = []
    try:
        for k in range(lib().sats_results_blocks(h)):
            flags = (C.c_int * 3)()
            lib().sats_results_flags(h, k, flags)
            n = lib().sats_results_rows(h, k)
            blk = dict(query=lib().sats_results_query(h, k).decode().strip(), dbfile=lib().sats_results_dbfile(h, k).decode().strip(),
                       ltype=bool(flags[0]), lorder=bool(flags[1]), lsoln=bool(flags[2]), names=[],
                       scores=np.zeros(n, np.int32), norm2=np.zeros(n), z=np.zeros(n), p=np.zeros(n), maps=[])
            name = C.create_string_buffer(9)
            sc = C.c_int32(); a = C.c_double(); z = C.c_double(); p = C.c_double()
            cap = MAXDIM_EXT
            pairs = np.zeros(2 * cap, np.int32)
            for r in range(n):
                lib().sats_results_row(h, k, r, name, C.byref(sc), C.byref(a), C.byref(z), C.byref(p))
                blk["names"].append(name.value.decode())
                blk["scores"][r], blk["norm2"][r], blk["z"][r], blk["p"][r] = sc.value, a.value, z.value, p.value
                m = min(max(lib().sats_results_map(h, k, r, pairs.ctypes.data, cap), 0), cap)     # malformed input may list more
                blk["maps"].append(pairs[:2 * m].reshape(m, 2).copy())
            out.append(blk)
    finally:
        lib().sats_results_free(h)
    return out


def roc_auc(scores, positive) -> float:
    s = np.ascontiguousarray(scores, np.float64)
    y = np.ascontiguousarray(positive, np.uint8)
    return lib().sats_roc_auc(s.ctypes.data, y.ctypes.data, len(s))
