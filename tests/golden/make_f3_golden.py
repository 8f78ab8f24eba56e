#!/usr/bin/env python3
"""Generates tests/golden/f3_golden.json: golden vectors for SURVEY 8(f3), query construction, produced by EXECUTING the
reference's own pure functions (read from /root/reference at generation time; nothing of them is copied into this repo):

  angle_to_tabcode     scripts/pttableau.py:434-469   angle -> two-letter tableau code (double quadrant encoding)
  LineLineIntersect    scripts/geometry.py:18-79      common perpendicular of two lines
  ProjectPointOntoLine scripts/geometry.py:82-110     foot of the perpendicular from a point to a line
  relative_angle       scripts/ptnode.py:752-880      interaxial angle omega of two SSE axes
  fit_axis             scripts/ptnode.py:1113-1292 (helix), :1846-2003 (strand)   SSE axis from the C-alpha trace

The scripts are python 2 and import Bio.PDB / Numeric, which are not installed here, so the three function bodies are cut
out of their files as text and executed in a namespace that supplies `pi`, numpy's `alltrue/less/abs`, `acos` and a minimal
stand-in for Bio.PDB.Vector with that class's documented operator semantics (`-`, `+ array`, `*` = dot, `**` = cross,
`/ scalar`, `norm`, `normsq`, `normalized`, `angle`, `[i]`, `get_array`) and numpy's SVD for `singular_value_decomposition`.  A structure built from random axes with those functions is written with the
database writer's conventions (scripts/convdb2.py:213-231: "%6s %4d", codes + ' ', "%6.3f " distances, NaN -> 0.000).

  python tests/golden/make_f3_golden.py          # needs /root/reference; rewrites f3_golden.json
"""
import json
import math
import re
import sys
from pathlib import Path

import numpy as np

REF = Path("/root/reference/scripts")
OUT = Path(__file__).resolve().parent / "f3_golden.json"


class Vector:
    """Bio.PDB.Vector's operator semantics, as far as the three functions use them."""

    def __init__(self, x, y=None, z=None):
        self._ar = np.array(x if y is None else (x, y, z), "d")

    def __sub__(self, other):
        return Vector(self._ar - (other._ar if isinstance(other, Vector) else np.array(other)))

    def __add__(self, other):
        return Vector(self._ar + (other._ar if isinstance(other, Vector) else np.array(other)))

    def __mul__(self, other):
        return float(sum(self._ar * other._ar))

    def __pow__(self, other):
        if isinstance(other, Vector):
            a, b = self._ar, other._ar
            return Vector(np.linalg.det(np.array(((a[1], a[2]), (b[1], b[2])))),
                          np.linalg.det(np.array(((a[2], a[0]), (b[2], b[0])))),
                          np.linalg.det(np.array(((a[0], a[1]), (b[0], b[1])))))
        return Vector(self._ar * np.array(other))

    def norm(self):
        return math.sqrt(sum(self._ar * self._ar))

    def normsq(self):
        return abs(sum(self._ar * self._ar))

    def __truediv__(self, x):
        return Vector(self._ar / np.array(x))

    __div__ = __truediv__

    def angle(self, other):
        c = (self * other) / (self.norm() * other.norm())
        c = min(c, 1)
        c = max(-1, c)
        return math.acos(c)

    def normalized(self):
        return Vector(self._ar / self.norm())

    def __getitem__(self, i):
        return self._ar[i]

    def get_array(self):
        return np.array(self._ar)


def cut(path, start_pat, end_pat, dedent=0):
    lines = path.read_text().split("\n")
    a = next(i for i, l in enumerate(lines) if re.match(start_pat, l))
    b = next(i for i in range(a + 1, len(lines)) if re.match(end_pat, lines[i]))
    pad = " " * dedent
    return "\n".join(l[dedent:] if l.startswith(pad) else l for l in lines[a:b + 1])     # column-0 comment lines stay comments


ns = {"pi": math.pi, "acos": math.acos, "alltrue": np.all, "less": np.less, "abs": np.abs, "Vector": Vector, "ALPHA": 100,
      "EPSILON": 1e-4, "verbose": False, "sys": sys, "min": min, "max": max, "array": np.array,
      "singular_value_decomposition": np.linalg.svd}
exec(cut(REF / "pttableau.py", r"^def angle_to_tabcode", r"^    return tabcode"), ns)
exec(cut(REF / "geometry.py", r"^def LineLineIntersect", r"^    return \(pa, pb, mua, mub\)"), ns)
exec(cut(REF / "ptnode.py", r"^    def relative_angle", r"^        return omega", dedent=4), ns)
angle_to_tabcode, relative_angle = ns["angle_to_tabcode"], ns["relative_angle"]
exec(cut(REF / "geometry.py", r"^def ProjectPointOntoLine", r"^    return Q"), ns)
helix_src = cut(REF / "ptnode.py", r"^    def fit_axis", r"^        return \(dircos, centroid\)", dedent=4)
exec(helix_src.replace("def fit_axis", "def fit_axis_helix"), ns)
_lines = (REF / "ptnode.py").read_text().split("\n")
_a = [i for i, l in enumerate(_lines) if re.match(r"^    def fit_axis", l)][1]                 # the second one: PTNodeStrand
_b = [i for i in range(_a, len(_lines)) if re.match(r"^            return \(dircos, centroid\)", _lines[i])][1]
strand_src = "\n".join(l[4:] if l.startswith("    ") else l for l in _lines[_a:_b + 1])
exec(strand_src.replace("def fit_axis", "def fit_axis_strand"), ns)


class Atom:
    def __init__(self, xyz):
        self.v = Vector(xyz)

    def get_vector(self):
        return self.v


class SSE:
    """Stands in for PTNodeHelix / PTNodeStrand: residues with a 'CA' atom, and the memo fields fit_axis uses."""

    def __init__(self, sse_type, ca):
        self.sse_type, self.ca = sse_type, np.asarray(ca, float).reshape(-1, 3)
        self.axis_direction_cosines = None
        self.axis_centroid = None
        self.seqnum = 0
        self.nodeid = "sse"

    def get_residue_list(self):
        return [{"CA": Atom(x)} for x in self.ca]

    def fit_axis(self, pdb_struct, mfile_fh=None):
        return (ns["fit_axis_strand"] if self.sse_type == 0 else ns["fit_axis_helix"])(self, pdb_struct, mfile_fh)

    def __str__(self):
        return "SSE(type %d, %d residues)" % (self.sse_type, len(self.ca))


def random_rotation(rng):
    q, _ = np.linalg.qr(rng.normal(size=(3, 3)))
    return q * np.sign(np.linalg.det(q))


def helix_trace(rng, n):
    t = np.arange(n)
    pts = np.stack([2.3 * np.cos(np.deg2rad(100.0) * t), 2.3 * np.sin(np.deg2rad(100.0) * t), 1.5 * t], 1)
    return (pts + rng.normal(scale=0.15, size=pts.shape)) @ random_rotation(rng).T + rng.uniform(-30, 30, 3)


def strand_trace(rng, n):
    t = np.arange(n)
    pts = np.stack([0.9 * (-1.0) ** t, np.zeros(n), 3.3 * t], 1)
    return (pts + rng.normal(scale=0.2, size=pts.shape)) @ random_rotation(rng).T + rng.uniform(-30, 30, 3)


class Axis:
    def __init__(self, centroid, dircos):
        self.c, self.d = Vector(centroid), Vector(dircos)

    def fit_axis(self, pdb_struct):
        return (self.d, self.c)


def tabcode(omega):
    try:
        return angle_to_tabcode(omega)
    except ValueError:
        return "ValueError"


def main():
    rng = np.random.default_rng(20240611)
    pi = math.pi
    # ---- angle -> code: every interval edge with its neighbours, a dense sweep, out-of-range values
    edges = [-pi, -3 * pi / 4, -pi / 2, -pi / 4, 0.0, pi / 4, pi / 2, 3 * pi / 4, pi]
    angles = []
    for e in edges:
        angles += [e, math.nextafter(e, -10.0), math.nextafter(e, 10.0), e - 1e-9, e + 1e-9]
    angles += list(np.linspace(-pi, pi, 721))
    angles += [float(x) for x in rng.uniform(-pi, pi, 500)]
    angles += [-0.0, 3.2, -3.2, 10.0, -10.0, float("inf"), float("-inf"), float("nan")]
    codes = [{"omega": float(a).hex() if not math.isnan(a) else "nan", "code": tabcode(a)} for a in angles]

    # ---- pairs of axes -> omega (self = first axis, SSE1 = second, as compute_tableau calls it for i < j)
    pairs = []

    def add_pair(c1, d1, c2, d2):
        om = relative_angle(Axis(c1, d1), Axis(c2, d2), None)
        pairs.append({"c1": list(map(float, c1)), "d1": list(map(float, d1)), "c2": list(map(float, c2)), "d2": list(map(float, d2)),
                      "omega": None if om is None else float(om).hex()})

    for _ in range(400):
        d1 = rng.normal(size=3); d1 /= np.linalg.norm(d1)
        d2 = rng.normal(size=3); d2 /= np.linalg.norm(d2)
        add_pair(rng.uniform(-30, 30, 3), d1, rng.uniform(-30, 30, 3), d2)
    add_pair([0, 0, 0], [1, 0, 0], [0, 5, 0], [1, 0, 0])            # parallel axes: no common perpendicular -> None
    add_pair([0, 0, 0], [1, 0, 0], [0, 5, 0], [-1, 0, 0])           # antiparallel: None
    add_pair([0, 0, 0], [1, 0, 0], [0, 0, 7], [0, 1, 0])            # orthogonal, skew
    add_pair([0, 0, 0], [1, 0, 0], [0, 0, -7], [0, 1, 0])
    add_pair([1, 2, 3], [0, 0, 1], [4, 5, 6], [0, 1, 1e-3])

    # ---- one whole structure: 14 SSEs with random axes, two of them parallel (a '??' entry), mixed types
    n = 14
    types = [0, 1, 0, 0, 3, 1, 0, 2, 1, 0, 0, 1, 3, 0]
    cent = rng.uniform(-25, 25, (n, 3))
    dirs = rng.normal(size=(n, 3))
    dirs /= np.linalg.norm(dirs, axis=1)[:, None]
    dirs[9] = dirs[2]                                                # parallel pair (2, 9)
    cent[13] = cent[0] + np.array([120.0, 0.0, 0.0])                 # a distance > 99.9 A
    tcode = ["e ", "xa", "xi", "xg"]
    rows = ["%6s %4d" % ("f3gold", n)]
    tab = [[None] * n for _ in range(n)]
    for i in range(n):
        for j in range(i + 1, n):
            om = relative_angle(Axis(cent[i], dirs[i]), Axis(cent[j], dirs[j]), None)
            tab[i][j] = tab[j][i] = "??" if om is None else tabcode(om)
        tab[i][i] = tcode[types[i]]
    for i in range(n):
        rows.append("".join(tab[i][j] + " " for j in range(i + 1)))
    for i in range(n):
        cells = []
        for j in range(i + 1):
            if i == j:
                dist = float(types[i])
            else:
                diff = cent[i] - cent[j]
                dist = float(np.sqrt(np.sum(diff * diff)))            # calc_sse_sse_midpoint_dist, ptdistmatrix.py:1009-1010
            if dist > 99.9:                                          # pytableaucreate.py:114-116 (the query writer's clamp)
                dist = 99.9
            cells.append("%6.3f " % dist)
        rows.append("".join(cells))
    structure = {"name": "f3gold", "n": n, "types": types, "centroid": cent.tolist(), "dircos": dirs.tolist(),
                 "ascii": "\n".join(rows) + "\n"}
    # ---- axes from C-alpha traces: helices and strands of every length from 1 (no axis) to 20 residues
    axes = []
    for sse_type, make in ((1, helix_trace), (3, helix_trace), (0, strand_trace)):
        for nres in list(range(1, 13)) + [15, 20]:
            for _ in range(2):
                ca = make(rng, nres)
                got = SSE(sse_type, ca).fit_axis(None)
                axes.append({"type": sse_type, "ca": ca.tolist(),
                             "dircos": None if got is None else [float(x) for x in got[0].get_array()],
                             "centroid": None if got is None else [float(x) for x in got[1].get_array()]})

    # ---- a structure straight from traces: nine SSEs, one of them a two-residue helix (no axis: '??' codes, 0.000 distances)
    ca_types = [0, 1, 0, 3, 1, 0, 0, 1, 2]
    ca_len = [6, 12, 5, 4, 2, 7, 3, 9, 5]
    traces = [(strand_trace if t == 0 else helix_trace)(rng, m) for t, m in zip(ca_types, ca_len)]
    nodes = [SSE(t, tr) for t, tr in zip(ca_types, traces)]
    m = len(nodes)
    rows = ["%6s %4d" % ("f3trac", m)]
    tab = [[None] * m for _ in range(m)]
    for i in range(m):
        for j in range(i + 1, m):
            if nodes[i].fit_axis(None) is None or nodes[j].fit_axis(None) is None:
                om = None
            else:
                om = relative_angle(Axis(nodes[i].axis_centroid.get_array(), nodes[i].axis_direction_cosines.get_array()),
                                    Axis(nodes[j].axis_centroid.get_array(), nodes[j].axis_direction_cosines.get_array()), None)
            tab[i][j] = tab[j][i] = "??" if om is None else tabcode(om)
        tab[i][i] = tcode[ca_types[i]]
    for i in range(m):
        rows.append("".join(tab[i][j] + " " for j in range(i + 1)))
    for i in range(m):
        cells = []
        for j in range(i + 1):
            if i == j:
                dist = float(ca_types[i])
            elif nodes[i].fit_axis(None) is None or nodes[j].fit_axis(None) is None:
                dist = 0.0                                             # None -> NaN in the matrix -> written as 0.000
            else:
                diff = nodes[i].axis_centroid.get_array() - nodes[j].axis_centroid.get_array()
                dist = float(np.sqrt(np.sum(diff * diff)))
            cells.append("%6.3f " % min(dist, 99.9))
        rows.append("".join(cells))
    from_ca = {"name": "f3trac", "types": ca_types, "traces": [t.tolist() for t in traces], "ascii": "\n".join(rows) + "\n"}
    OUT.write_text(json.dumps({"generator": "tests/golden/make_f3_golden.py", "codes": codes, "pairs": pairs,
                               "structure": structure, "axes": axes, "from_ca": from_ca}, separators=(",", ":")))
    print("wrote", OUT, len(codes), "angles,", len(pairs), "axis pairs,", len(axes), "fitted axes")


if __name__ == "__main__":
    main()
