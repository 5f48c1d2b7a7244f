"""rfluxmtx front-end on top of the CUDA rcontrib path.

Restates /root/reference/src/radiance/util/rfluxmtx.c: the `#@rfluxmtx h=.. u=..
o=..` directives of receiver and sender files, the rcontrib command line it
synthesises (`finish_receiver`, :427-583), pass-through mode (sender `-`, rays
on stdin) and sampling mode (sender surface: stratified origins and directions
per sender bin, `sample_uniform/_shirchiu/_reinhart/_klems`, :771-926).  The
reference then spawns `rcontrib` with `!oconv -f ... receiver` as its octree;
here the same command line goes to `rt.rcontrib_main` in-process and the octree
comes from the library's own builder (`rb_oconv_files`).

Same Python signature as `pyradiance.rfluxmtx` (src/pyradiance/util.py:833-870).

Differences, by design: `!command` inputs are not executed; the 1-D -> n-D
sample spreading uses bit de-interleaving where the reference walks a Hilbert
curve (`SDmultiSamp`, common/bsdf.c:531-557) -- both are measure-preserving, so
sender sampling agrees in distribution, not sample for sample.
"""
from __future__ import annotations

import math
import os
import re
import tempfile
from pathlib import Path
from typing import Sequence

import numpy as np

from . import _lib
from ._lib import RBError

PARAMSTART = "@rfluxmtx"
_ST_POLY, _ST_RING, _ST_SOURCE = 1, 2, 3
_SURF_TYPES = {"polygon": _ST_POLY, "ring": _ST_RING, "source": _ST_SOURCE}
FTINY = 1e-6


def _g(v: float) -> str:
    """printf("%g")"""
    return "%g" % v


# ------------------------------------------------------------------ parsing --
class _Surf:
    def __init__(self, styp, name, farg):
        self.styp, self.name, self.farg = styp, name, np.asarray(farg, dtype=float)
        self.snrm = np.zeros(3)
        self.area = 0.0
        self.tris = None          # polygon: list of (afrac, (a, b, c))
        self.uva = None           # ring: tangent axes


class _Params:
    def __init__(self):
        self.sign, self.hemis, self.hsiz = "+", "", 0
        self.slist: list[_Surf] = []
        self.vup = np.zeros(3)
        self.nrm = np.zeros(3)
        self.outfn = None

    def reset(self):
        self.slist, self.vup, self.nrm, self.outfn = [], np.zeros(3), np.zeros(3), None


def _get_direction(s: str):
    """rfluxmtx.c:296-333"""
    sign, k = 1.0, 0
    while k < len(s) and s[k] in "+-":
        if s[k] == "-":
            sign = -sign
        k += 1
    if k < len(s) and s[k] in "xyzXYZ":
        if k + 1 < len(s) and not s[k + 1].isspace():
            return None
        dv = np.zeros(3)
        dv["xyz".index(s[k].lower())] = sign
        return dv
    try:
        parts = [float(x) for x in s[k:].split()[0].split(",")]
    except (ValueError, IndexError):
        return None
    if len(parts) != 3:
        return None
    dv = np.array(parts)
    dv[0] *= sign
    n = np.linalg.norm(dv)
    return dv / n if n > 0 else None


def _parse_params(p: _Params, text: str):
    """rfluxmtx.c:336-412: h=[+-]type, u=dir, o=file"""
    pos = 0
    while pos < len(text):
        c = text[pos]
        pos += 1
        if c in " \t\r\n":
            continue
        if pos >= len(text) or text[pos] != "=":
            raise RBError(f"rfluxmtx: bad parameter string:{text}")
        pos += 1
        if c == "h":
            if pos < len(text) and text[pos] in "+-":
                p.sign = text[pos]
                pos += 1
            else:
                p.sign = "+"
            m = re.match(r"\S+", text[pos:])
            if not m:
                raise RBError(f"rfluxmtx: bad parameter string:{text}")
            p.hemis = m.group(0)
            digits = "".join(ch for ch in p.hemis if ch.isdigit())
            p.hsiz = int(digits) if digits else 0
            p.hsiz += not p.hsiz
            pos += m.end()
        elif c == "u":
            m = re.match(r"\S+", text[pos:])
            dv = _get_direction(text[pos:]) if m else None
            if dv is None:
                raise RBError(f"rfluxmtx: bad parameter string:{text}")
            p.vup = dv
            pos += m.end()
        elif c == "o":
            if pos < len(text) and text[pos] in "\"'":
                q = text[pos]
                end = text.find(q, pos + 1)
                if end < 0:
                    raise RBError(f"rfluxmtx: bad parameter string:{text}")
                p.outfn = text[pos + 1:end]
                pos = end + 1
            else:
                m = re.match(r"\S+", text[pos:])
                if not m:
                    raise RBError(f"rfluxmtx: bad parameter string:{text}")
                p.outfn = m.group(0)
                pos += m.end()
        else:
            raise RBError(f"rfluxmtx: bad parameter string:{text}")


def _scan_scene(path):
    """load_scene(), rfluxmtx.c:1198-1251: yields ('params', text) for `#@rfluxmtx`
    lines and ('object', mod, type, name, sargs, fargs) for scene objects."""
    path = str(path)
    if path.startswith("!"):
        raise RBError(f"rfluxmtx: input from command '{path}' is not supported (commands are not executed)")
    try:
        text = Path(path).read_text()
    except OSError:
        raise RBError(f"rfluxmtx: cannot load '{path}'")
    toks = []            # tokens with ('P', text) markers
    for line in text.splitlines():
        s = line.strip()
        if not s:
            continue
        if s.startswith("!"):
            raise RBError(f"rfluxmtx: ({path}) '!command' lines are not supported (commands are not executed)")
        if "#" in line:
            head, _, tail = line.partition("#")
            toks.extend(head.split())
            if tail.startswith(PARAMSTART) and (len(tail) == len(PARAMSTART) or tail[len(PARAMSTART)].isspace()):
                toks.append(("P", tail[len(PARAMSTART):] + "\n"))
            continue
        toks.extend(line.split())
    i = 0
    while i < len(toks):
        t = toks[i]
        if isinstance(t, tuple):
            yield ("params", t[1])
            i += 1
            continue
        # an object: mod type name, then three counted argument lists
        def nxt():
            nonlocal i
            while i < len(toks) and isinstance(toks[i], tuple):
                i += 1                      # (a directive inside an object is dropped, as fscanf would choke on it)
            if i >= len(toks):
                raise RBError(f"rfluxmtx: ({path}) unexpected end of file")
            i += 1
            return toks[i - 1]
        mod, typ, name = nxt(), nxt(), nxt()
        if typ == "alias":
            nxt()
            yield ("object", mod, typ, name, [], [])
            continue
        ns = int(nxt()); sargs = [nxt() for _ in range(ns)]
        ni = int(nxt())
        for _ in range(ni):
            nxt()
        nf = int(nxt()); fargs = [float(nxt()) for _ in range(nf)]
        yield ("object", mod, typ, name, sargs, fargs)


def _make_surface(st, name, farg) -> _Surf | None:
    """add_surface(), rfluxmtx.c:1035-1113"""
    s = _Surf(st, name, farg)
    n = len(farg)
    if st == _ST_RING:
        if n != 8:
            raise RBError(f"rfluxmtx: bad argument count for surface element '{name}'")
        nl = np.linalg.norm(s.farg[3:6])
        if nl == 0:
            raise RBError(f"rfluxmtx: bad orientation for surface element '{name}'")
        s.snrm = s.farg[3:6] / nl
        if s.farg[7] < s.farg[6]:
            s.farg[6], s.farg[7] = s.farg[7], s.farg[6]
        s.area = math.pi * (s.farg[7] ** 2 - s.farg[6] ** 2)
    elif st == _ST_POLY:
        if n < 9 or n % 3:
            raise RBError(f"rfluxmtx: bad argument count for surface element '{name}'")
        v = s.farg.reshape(-1, 3)
        nsum = np.zeros(3)
        e1 = v[1] - v[0]
        for k in range(2, len(v)):
            e2 = v[k] - v[0]
            nsum += np.cross(e1, e2)
            e1 = e2
        ln = np.linalg.norm(nsum)
        s.area = ln * 0.5
        s.snrm = nsum / ln if ln > 0 else nsum
    else:
        if n != 4:
            raise RBError(f"rfluxmtx: bad argument count for surface element '{name}'")
        nl = np.linalg.norm(s.farg[:3])
        if nl == 0:
            raise RBError(f"rfluxmtx: bad orientation for surface element '{name}'")
        s.snrm = -s.farg[:3] / nl                        # need to reverse "normal"
        a = math.sin((math.pi / 180. / 2.) * s.farg[3])
        s.area = math.pi * a * a
    if s.area <= FTINY * FTINY:
        return None                                       # (warning - zero area)
    return s


def _finish_receiver(p: _Params, mod: str, rcarg: list, used: set, binjitter):
    """finish_receiver(), rfluxmtx.c:427-583"""
    if not mod:
        raise RBError("rfluxmtx: missing receiver surface!")
    if p.outfn is not None:
        rcarg += ["-o", p.outfn]
    if not p.hemis:
        raise RBError("rfluxmtx: missing hemisphere sampling type!")
    nl = np.linalg.norm(p.nrm)
    if nl == 0:
        raise RBError("rfluxmtx: undefined normal for hemisphere sampling")
    nrm = p.nrm / nl
    vl = np.linalg.norm(p.vup)
    vup = p.vup / vl if vl > 0 else (np.array([0., 0., 1.]) if abs(nrm[2]) < .7 else np.array([0., 1., 0.]))
    h = p.hemis
    h0, h1 = h[0].lower(), (h[1].lower() if len(h) > 1 else "")
    calfn = params = binv = binf = nbins = None
    uniform = False
    jit = f",JTR={binjitter}" if binjitter is not None else ""
    six = lambda: f"rNx={_g(nrm[0])},rNy={_g(nrm[1])},rNz={_g(nrm[2])},Ux={_g(vup[0])},Uy={_g(vup[1])},Uz={_g(vup[2])},RHS={p.sign}1"

    def cal(name):
        if name in used:
            return None
        used.add(name)
        return name

    if h0 == "u" or h[0] == "1":
        if p.slist[0].styp != _ST_SOURCE:
            binv = f"if(-Dx*{_g(nrm[0])}-Dy*{_g(nrm[1])}-Dz*{_g(nrm[2])},0,-1)"
        else:
            binv = "0"
        nbins = "1"
        uniform = True
    elif h0 == "s" and h1 == "c":
        if p.hsiz <= 1:
            raise RBError("rfluxmtx: missing size for Shirley-Chiu sampling!")
        calfn = cal("disk2square.cal")
        params = f"SCdim={p.hsiz}," + six() + jit
        binv, nbins = "scbin", "SCdim*SCdim"
    elif h0 in "rt":
        calfn = cal("reinhartb.cal")
        params = f"MF={p.hsiz}," + six() + jit
        binv, nbins = "rbin", "Nrbins"
    elif h0 == "k" and (len(h) == 1 or h1 == "f" or h[1] == "1"):
        calfn, binf, nbins = cal("klems_full.cal"), "kbin", "Nkbins"
    elif h0 == "k" and (h1 == "h" or h[1] == "2"):
        calfn, binf, nbins = cal("klems_half.cal"), "khbin", "Nkhbins"
    elif h0 == "k" and (h1 == "q" or h[1] == "4"):
        calfn, binf, nbins = cal("klems_quarter.cal"), "kqbin", "Nkqbins"
    elif h.lower() == "cie":
        raise RBError("rfluxmtx: h=cie (cieskyscan.cal) is not built as a native bin function")
    else:
        raise RBError(f"rfluxmtx: unrecognized hemisphere sampling: h={h}")
    if h0 == "k":
        params = f"RHS={p.sign}1" + jit
    if not uniform:
        for sp in p.slist:
            if sp.styp == _ST_SOURCE and abs(sp.area - math.pi) > 1e-3:
                raise RBError(f"rfluxmtx: source '{sp.name}' must be 180-degrees")
    if calfn is not None:
        rcarg += ["-f", calfn]
    if params is not None:
        rcarg += ["-p", params]
    if nbins is not None:
        rcarg += ["-bn", nbins]
    if binv is not None:
        rcarg += ["-b", binv]
    elif binf is not None:
        rcarg += ["-b", f"{binf}({_g(nrm[0])},{_g(nrm[1])},{_g(nrm[2])},{_g(vup[0])},{_g(vup[1])},{_g(vup[2])})"]
    rcarg += ["-m", mod]


def _load_receivers(path, rcarg, binjitter):
    """load_scene(.., add_recv_object) + finish_receiver(), rfluxmtx.c:1116-1151"""
    cur = _Params()
    curmod, newparams, used = "", "", set()
    for item in _scan_scene(path):
        if item[0] == "params":
            newparams += item[1]
            continue
        _, mod, typ, name, sargs, fargs = item
        st = _SURF_TYPES.get(typ)
        if st is None:
            continue
        if mod != curmod:
            if curmod:
                _finish_receiver(cur, curmod, rcarg, used, binjitter)
                cur.reset()
            _parse_params(cur, newparams)
            newparams = ""
            curmod = mod
        s = _make_surface(st, name, fargs)
        if s is not None:
            cur.nrm = cur.nrm + s.snrm * s.area
            cur.slist.insert(0, s)
    _finish_receiver(cur, curmod, rcarg, used, binjitter)


def _load_sender(path) -> _Params:
    """load_scene(.., add_send_object), rfluxmtx.c:1154-1195"""
    cur = _Params()
    newparams = ""
    for item in _scan_scene(path):
        if item[0] == "params":
            newparams += item[1]
            continue
        _, mod, typ, name, sargs, fargs = item
        st = _SURF_TYPES.get(typ)
        if st is None:
            continue
        if st == _ST_SOURCE:
            raise RBError("rfluxmtx: cannot use source as a sender!")
        _parse_params(cur, newparams)
        newparams = ""
        s = _make_surface(st, name, fargs)
        if s is not None:
            cur.nrm = cur.nrm + s.snrm * s.area
            cur.slist.insert(0, s)
    return cur


# ----------------------------------------------------------------- sampling --
_KLEMS = {"1": ([0., 5., 15., 25., 35., 45., 55., 65., 75., 90.], [1, 8, 16, 20, 24, 24, 24, 16, 12]),
          "2": ([0., 6.5, 19.5, 32.5, 45.5, 58.5, 71.5, 90.], [1, 8, 12, 16, 20, 12, 8]),
          "4": ([0., 9., 27., 45., 63., 90.], [1, 8, 12, 12, 8])}        # common/bsdf_m.c:31-63
_TNAZ = (30, 30, 24, 24, 18, 12, 6)


def _multisamp(x: np.ndarray, n: int, rng) -> np.ndarray:
    """One uniform variable -> n stratified ones (role of SDmultiSamp, bsdf.c:531-557):
    the bits of x are dealt round-robin to the n coordinates, the rest is jitter."""
    x = np.clip(x, 0.0, 0.999999999999999)
    if n == 1:
        return x[:, None]
    nbits = 48 // n
    ndx = (x * float(1 << (nbits * n))).astype(np.uint64)
    coord = np.zeros((x.shape[0], n), dtype=np.uint64)
    for b in range(nbits):
        for k in range(n):
            bit = (ndx >> np.uint64(b * n + k)) & np.uint64(1)
            coord[:, k] |= bit << np.uint64(b)
    return (coord.astype(np.float64) + rng.random((x.shape[0], n))) / float(1 << nbits)


def _square2disk(sx, sy):
    """common/disk2square.c:43-79"""
    a, b = 2. * sx - 1., 2. * sy - 1.
    r = np.zeros_like(a)
    phi = np.zeros_like(a)
    with np.errstate(divide="ignore", invalid="ignore"):
        m1 = (a > -b) & (a > b);  r[m1] = a[m1];  phi[m1] = (math.pi / 4) * (b[m1] / a[m1])
        m2 = (a > -b) & ~(a > b); r[m2] = b[m2];  phi[m2] = (math.pi / 4) * (2. - a[m2] / b[m2])
        m3 = ~(a > -b) & (a < b); r[m3] = -a[m3]; phi[m3] = (math.pi / 4) * (4. + b[m3] / a[m3])
        m4 = ~(a > -b) & ~(a < b); r[m4] = -b[m4]
        phi[m4] = np.where(b[m4] != 0., (math.pi / 4) * (6. - a[m4] / np.where(b[m4] != 0., b[m4], 1.)), 0.)
    r = r * 0.9999999999999
    return r * np.cos(phi), r * np.sin(phi)


def _perp(n, rng):
    """any unit vector perpendicular to n (make_axes(), rfluxmtx.c:586-596)"""
    for _ in range(64):
        v = rng.normal(size=3)
        u = np.cross(v, n)
        ln = np.linalg.norm(u)
        if ln > 1e-3:
            return u / ln
    raise RBError("rfluxmtx: bad surface normal in make_axes!")


def _triangulate(s: _Surf, rng):
    """ssamp_poly() set-up, rfluxmtx.c:651-700: ear clipping in the polygon's plane"""
    v = s.farg.reshape(-1, 3)
    if len(v) == 3:
        s.tris = [(1.0, (0, 1, 2))]
        return
    u0 = _perp(s.snrm, rng)
    u1 = np.cross(s.snrm, u0)
    p2 = np.stack([v @ u0, v @ u1], axis=1)
    idx = list(range(len(v)))
    tris = []

    def area2(a, b, c):
        return (p2[b, 0] - p2[a, 0]) * (p2[c, 1] - p2[a, 1]) - (p2[c, 0] - p2[a, 0]) * (p2[b, 1] - p2[a, 1])

    def inside(q, a, b, c):
        d1, d2, d3 = area2(a, b, q), area2(b, c, q), area2(c, a, q)
        return (d1 > 0 and d2 > 0 and d3 > 0)

    guard = 0
    while len(idx) > 3 and guard < 10000:
        guard += 1
        clipped = False
        for k in range(len(idx)):
            a, b, c = idx[k - 1], idx[k], idx[(k + 1) % len(idx)]
            if area2(a, b, c) <= 0:
                continue                                  # reflex corner
            if any(inside(q, a, b, c) for q in idx if q not in (a, b, c)):
                continue
            tris.append((a, b, c))
            idx.pop(k)
            clipped = True
            break
        if not clipped:
            raise RBError(f"rfluxmtx: cannot triangulate polygon '{s.name}'")
    tris.append(tuple(idx))
    out = []
    for a, b, c in tris:
        out.append((area2(a, b, c) / (2. * s.area), (a, b, c)))
    s.tris = out


def _samp_surface(s: _Surf, x: np.ndarray, rng) -> np.ndarray:
    """ssamp_poly / ssamp_ring: origins on one surface from uniform x"""
    if s.styp == _ST_RING:
        if s.uva is None:
            u0 = _perp(s.snrm, rng)
            s.uva = (u0, np.cross(s.snrm, u0))
        s2 = _multisamp(x, 2, rng)
        r = np.sqrt(s2[:, 0] * s.area * (1. / math.pi) + s.farg[6] ** 2)
        ang = s2[:, 1] * 2. * math.pi
        return s.farg[:3] + (r * np.cos(ang))[:, None] * s.uva[0] + (r * np.sin(ang))[:, None] * s.uva[1]
    if s.tris is None:
        _triangulate(s, rng)
    v = s.farg.reshape(-1, 3)
    out = np.zeros((x.shape[0], 3))
    rem = x.copy()
    done = np.zeros(x.shape[0], dtype=bool)
    for k, (afrac, (a, b, c)) in enumerate(s.tris):
        last = k == len(s.tris) - 1
        pick = ~done & ((rem <= afrac) | last)
        if pick.any():
            s2 = _multisamp(rem[pick] / afrac, 2, rng)
            t1 = np.sqrt(s2[:, 1])
            t0 = s2[:, 0] * t1
            t1 = 1. - t1
            out[pick] = v[a] + t0[:, None] * (v[b] - v[a]) + t1[:, None] * (v[c] - v[a])
            done |= pick
        rem = rem - afrac
    return out


def _sample_origin(p: _Params, rdir: np.ndarray, x: np.ndarray, rng) -> np.ndarray:
    """sample_origin(), rfluxmtx.c:724-768"""
    if len(p.slist) == 1:
        sp = p.slist[0]
        if np.any(rdir @ sp.snrm >= FTINY):
            raise RBError(f"rfluxmtx: internal - sample behind sender '{sp.name}'")
        return _samp_surface(sp, x, rng)
    proj = np.stack([np.maximum(-(rdir @ sp.snrm) * sp.area, 0.) for sp in p.slist], axis=1)
    tarea = proj.sum(1)
    if np.any(tarea < FTINY * FTINY):
        raise RBError("rfluxmtx: internal - sample behind all sender elements!")
    t = tarea * x
    out = np.zeros((x.shape[0], 3))
    done = np.zeros(x.shape[0], dtype=bool)
    for k, sp in enumerate(p.slist):
        last = k == len(p.slist) - 1
        pick = ~done & ((t <= proj[:, k]) | last)
        if pick.any():
            out[pick] = _samp_surface(sp, t[pick] / np.where(proj[pick, k] > 0, proj[pick, k], 1.), rng)
            done |= pick
        t = t - proj[:, k]
    return out


def _prepare_sampler(p: _Params):
    """prepare_sampler(), rfluxmtx.c:929-1009: returns (kind, nbins) and sets p.udir/p.vdir"""
    if not p.slist:
        raise RBError("rfluxmtx: no sender surface!")
    if not p.hemis:
        raise RBError("rfluxmtx: missing sender sampling type!")
    nl = np.linalg.norm(p.nrm)
    if nl == 0:
        raise RBError("rfluxmtx: undefined normal for sender sampling")
    p.nrm = p.nrm / nl
    vl = np.linalg.norm(p.vup)
    p.vup = p.vup / vl if vl > 0 else (np.array([0., 0., 1.]) if abs(p.nrm[2]) < .7 else np.array([0., 1., 0.]))
    u = np.cross(p.vup, p.nrm)
    ul = np.linalg.norm(u)
    if ul == 0:
        raise RBError("rfluxmtx: up vector coincides with sender normal")
    p.udir = u / ul
    p.vdir = np.cross(p.nrm, p.udir)
    if p.sign == "-":
        p.udir = -p.udir
    h = p.hemis
    h0 = h[0].lower()
    if h0 == "u" or h[0] == "1":
        return "u", 1
    if h0 == "s" and len(h) > 1 and h[1].lower() == "c":
        return "sc", p.hsiz * p.hsiz
    if h0 in "rt":
        rowmax = 7 * p.hsiz + 1
        return "r", sum((1 if r >= rowmax - 1 else p.hsiz * _TNAZ[r // p.hsiz]) for r in range(rowmax))
    if h0 == "k":
        k = {"": "1", "f": "1", "1": "1", "h": "2", "2": "2", "q": "4", "4": "4"}.get(h[1:2].lower())
        if k is None:
            raise RBError(f"rfluxmtx: unrecognized sender sampling: h={h}")
        p.kbasis = k
        return "k", sum(_KLEMS[k][1])
    raise RBError(f"rfluxmtx: unrecognized sender sampling: h={h}")


def _sample_sender(p: _Params, kind: str, nbins: int, sampcnt: int, rng) -> np.ndarray:
    """All sender rays, bin-major: [nbins * sampcnt, 6] (the stream rfluxmtx pipes into rcontrib)."""
    b = np.repeat(np.arange(nbins), sampcnt)
    n = np.tile(np.arange(sampcnt - 1, -1, -1), nbins)         # `while (n--)`
    x = (n + rng.random(n.shape[0])) / sampcnt
    if kind == "k":
        s2 = _multisamp(x, 2, rng)
        tmin, nphis = _KLEMS[p.kbasis]
        starts = np.concatenate([[0], np.cumsum(nphis)])
        li = np.searchsorted(starts, b, side="right") - 1
        ndx = b - starts[li]
        rx = _multisamp(s2[:, 1], 2, rng)                        # fo_getvec(): randX = fractional part
        c0 = np.cos(np.radians(np.asarray(tmin)[li])) ** 2
        c1 = np.cos(np.radians(np.asarray(tmin)[li + 1])) ** 2
        d = np.sqrt((1. - rx[:, 0]) * c0 + rx[:, 0] * c1)
        azi = 2. * math.pi * (ndx + rx[:, 1] - .5) / np.asarray(nphis)[li]
        sp = np.sqrt(1. - d * d)
        duvw = np.stack([np.cos(azi) * sp, np.sin(azi) * sp, d], axis=1)
        s0 = s2[:, 0]
        sgn = -1.
    else:
        s3 = _multisamp(x, 3, rng)
        s0 = s3[:, 0]
        if kind == "u":
            dx, dy = _square2disk(s3[:, 1], s3[:, 2])
            duvw = np.stack([dx, dy, -np.sqrt(1. - dx * dx - dy * dy)], axis=1)
            sgn = 1.
        elif kind == "sc":
            dx, dy = _square2disk((b // p.hsiz + s3[:, 1]) / p.hsiz, (b % p.hsiz + s3[:, 2]) / p.hsiz)
            duvw = np.stack([dx, dy, np.sqrt(1. - dx * dx - dy * dy)], axis=1)
            sgn = -1.
        else:
            rowmax = 7 * p.hsiz + 1
            rnaz = np.array([(1 if r >= rowmax - 1 else p.hsiz * _TNAZ[r // p.hsiz]) for r in range(rowmax)])
            starts = np.concatenate([[0], np.cumsum(rnaz)])
            row = np.searchsorted(starts, b, side="right") - 1
            col = b - starts[row]
            rah = (.5 * math.pi) / (rowmax - .5)
            s1 = np.where(row >= rowmax - 1, s3[:, 1] ** 2, s3[:, 1])       # avoid crowding at zenith
            alt = (row + s1) * rah
            azi = (2. * math.pi) * (col + s3[:, 2] - .5) / rnaz[row]
            ca = np.cos(alt)
            duvw = np.stack([np.sin(azi) * ca, -np.cos(azi) * ca, np.sqrt(1. - ca * ca)], axis=1)
            sgn = -1.
    rdir = sgn * (duvw[:, 0:1] * p.udir + duvw[:, 1:2] * p.vdir + duvw[:, 2:3] * p.nrm)
    org = _sample_origin(p, rdir, s0, rng)
    return np.concatenate([org, rdir], axis=1)


# ------------------------------------------------------------- command line --
def rcontrib_command(argv: Sequence[str]):
    """Screen an rfluxmtx command line (argv[0] = program name) the way main() does
    (rfluxmtx.c:1254-1458) and return (rcontrib argv WITHOUT the octree, sender file or None,
    [receiver, scene inputs...], sample count, verbose).  The returned argv is what the
    reference prints with -v as `rcontrib ...` in front of its "!oconv ..." octree."""
    argv = [str(a) for a in argv]
    rc = ["rcontrib", "-fo+"]
    fmt = ["a", "a"]
    sampcnt = 0
    xrs = yrs = ldopt = iropt = binjitter = None
    verbose = False
    a = 1
    userr = RBError("Usage: rfluxmtx [-v][-bj frac][rcontrib options] sender.rad receiver.rad [-i system.oct] [system.rad ..]")
    while a < len(argv) - 2:
        s = argv[a]
        if not s.startswith("-") or len(s) < 2:
            break
        na = 1
        c = s[1]
        if c == "v":
            verbose = not verbose; a += 1; continue
        if c == "f":
            c2 = s[2:3]
            if c2 == "":
                na = 2
            elif c2 == "o":
                raise userr
            elif c2 in "afdc":
                fmt = [c2, s[3:4] or c2]; a += 1; continue
            else:
                raise userr
        elif c == "x":
            xrs = argv[a + 1]; a += 2; continue
        elif c == "y":
            yrs = argv[a + 1]; a += 2; continue
        elif c == "c":
            c2 = s[2:3]
            if c2 == "s":
                na = 2
            elif c2 == "w":
                na = 3
            elif c2 == "":
                sampcnt = int(argv[a + 1])
                if sampcnt <= 0:
                    raise userr
                a += 2; continue
        elif c in "Ii":
            iropt = s; a += 1; continue
        elif c == "w":
            pass
        elif c in "Vuhr":
            pass
        elif c in "nsote":
            na = 2
        elif c == "b":
            if s[2:3] == "j":
                binjitter = argv[a + 1]; a += 2; continue
            if s[2:3] != "v":
                raise userr
        elif c == "l":
            if s[2:3] == "d":
                ldopt = s; a += 1; continue
            na = 2
        elif c == "d":
            if s[2:3] != "v":
                na = 2
        elif c == "a":
            if s[2:3] == "p":
                raise RBError("rfluxmtx: photon maps (-ap) are not built")
            na = 4 if s[2:3] == "v" else 2
        elif c == "m":
            if not s[2:3]:
                raise userr
            na = 4 if s[2:3] in "ea" else 2
        else:
            raise RBError(f"rfluxmtx: unsupported option '{s}'")
        rc += argv[a:a + na]
        a += na
    if a > len(argv) - 2:
        raise userr
    sendfn = argv[a]; a += 1
    if sendfn.startswith("-"):
        if len(sendfn) > 1:
            raise userr
        sendfn = None
        if iropt:
            rc.append(iropt)
        if xrs:
            rc += ["-x", xrs]
        if yrs:
            rc += ["-y", yrs]
        if ldopt:
            rc.append(ldopt)
        if sampcnt <= 0:
            sampcnt = 1
    else:
        if iropt:
            raise RBError("rfluxmtx: -i, -I supported for pass-through only")
        fmt[0] = "d"
        if sampcnt <= 0:
            sampcnt = 10000
    rc += [f"-f{fmt[0]}{fmt[1]}", "-c", str(sampcnt)]
    _load_receivers(argv[a], rc, binjitter)
    return rc, sendfn, argv[a:], sampcnt, verbose


def _build_octree(inputs: Sequence[str], workdir: Path) -> Path:
    """oconv_command(), rfluxmtx.c:139-187: `oconv -f [-i octree] scene... receiver` (receiver goes last)."""
    recv, rest = inputs[0], list(inputs[1:])
    include = None
    files = []
    k = 0
    while k < len(rest):
        if rest[k] == "-i":
            if include is not None:
                raise RBError("rfluxmtx: only one -i octree can be included")
            include = rest[k + 1]
            k += 2
            continue
        if rest[k].startswith("-"):
            raise RBError(f"rfluxmtx: unsupported oconv option '{rest[k]}'")
        files.append(rest[k])
        k += 1
    files.append(recv)
    octf = workdir / "rfluxmtx_scene.oct"
    _lib.oconv_files(files, octf, include_octree=include)
    return octf


def rfluxmtx_main(argv: Sequence[str], stdin: bytes | None = None, device: int = 0, seed: int | None = None) -> bytes:
    """The rfluxmtx command: argv[0] is the program name."""
    from .rt import rcontrib_main
    rc, sendfn, inputs, sampcnt, verbose = rcontrib_command(argv)
    if any(",JTR=" in a for a in rc):
        raise RBError("rfluxmtx: bin jitter (-bj) is not built (the bin functions are native code, not .cal files)")
    with tempfile.TemporaryDirectory(prefix="rb200_rfluxmtx_") as td:
        octf = _build_octree(inputs, Path(td))
        if sendfn is None:                              # pass-through mode: rcontrib does everything
            return rcontrib_main(rc + [str(octf)], stdin or b"", device=device)
        rng = np.random.default_rng(seed)
        p = _load_sender(sendfn)
        kind, nsbins = _prepare_sampler(p)
        rays = _sample_sender(p, kind, nsbins, sampcnt, rng)
        return rcontrib_main(rc + ["-y", str(nsbins), str(octf)], np.ascontiguousarray(rays).tobytes(), device=device)


def rfluxmtx(receiver, surface=None, rays: bytes | None = None, params: Sequence[str] | None = None, octree=None,
             scene: Sequence | None = None) -> bytes:
    """Same call as pyradiance.rfluxmtx (src/pyradiance/util.py:833-870)."""
    cmd = ["rfluxmtx"]
    if params:
        cmd.extend(str(p) for p in params)
    cmd.append(os.fspath(surface) if surface is not None else "-")
    cmd.append(os.fspath(receiver))
    if octree is not None:
        cmd.extend(["-i", os.fspath(octree)])
    if scene is not None:
        cmd.extend(os.fspath(s) for s in scene)
    return rfluxmtx_main(cmd, rays)
