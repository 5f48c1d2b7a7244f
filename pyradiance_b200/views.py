"""vwrays: view rays for rtrace / rcontrib (the step before the hot path).

Restates /root/reference/src/radiance/util/vwrays.c (:40-170 main, :245-300 putrays,
:191-243 pix2rays) on top of common/image.c (`setview` :24-127, `normaspect` :197-211,
`viewray` :214-305, `pix2loc` :398-421, `getviewopt` :451-517).  Vectorised numpy on the
host: 2048 x 2048 rays take a fraction of a second, and the arithmetic follows the
reference's expression order so deterministic views reproduce its doubles.

Same Python signature as `pyradiance.vwrays` (src/pyradiance/util.py:1025-1066).
Not built: pictures / depth buffers as the view source (`pic`, `zbuf`) and the
depth-of-field aperture (`-pd`); pixel jitter (`-pj`) draws from numpy's generator,
so jittered rays agree with the reference in distribution only.
"""
from __future__ import annotations

import math
import re
from typing import Sequence

import numpy as np

from ._lib import RBError, format_ascii  # noqa: F401

FTINY = 1e-6


class View:
    """VIEW of common/view.h with STDVIEW defaults"""

    def __init__(self):
        self.type = "v"
        self.vp = np.zeros(3)
        self.vdir = np.array([0., 1., 0.])
        self.vup = np.array([0., 0., 1.])
        self.vdist, self.horiz, self.vert = 1., 45., 45.
        self.hoff = self.voff = self.vfore = self.vaft = 0.
        self.hvec = np.zeros(3)
        self.vvec = np.zeros(3)
        self.hn2 = self.vn2 = 0.


def _normalize(v):
    """common/fvect.c:130-157 (the near-unit shortcut included)"""
    d = float(v[0] * v[0] + v[1] * v[1] + v[2] * v[2])
    if d == 0.0:
        return 0.0, v
    if 1.0 - FTINY <= d <= 1.0 + FTINY:
        ln = 0.5 + 0.5 * d
        d = 2.0 - ln
    else:
        ln = math.sqrt(d)
        d = 1.0 / ln
    return ln, v * d


def _normalize_rows(v):
    d = v[:, 0] * v[:, 0] + v[:, 1] * v[:, 1] + v[:, 2] * v[:, 2]
    near = (d <= 1.0 + FTINY) & (d >= 1.0 - FTINY)
    with np.errstate(divide="ignore", invalid="ignore"):
        ln = np.where(near, 0.5 + 0.5 * d, np.sqrt(d))
        s = np.where(near, 2.0 - ln, 1.0 / ln)
    s = np.where(d == 0.0, 0.0, s)
    return np.where(d == 0.0, 0.0, ln), v * s[:, None]


def get_view_opts(v: View, av: Sequence[str]) -> int:
    """Consume one -v? option at av[0] (getviewopt); returns the number of extra words used."""
    a = av[0]
    if len(a) < 3 or a[:2] != "-v":
        raise RBError(f"vwrays: bad view option '{a}'")
    c = a[2]

    def f(k):
        try:
            return float(av[k])
        except (IndexError, ValueError):
            raise RBError(f"vwrays: bad arguments for '{a}'")
    if c == "t":
        if len(a) != 4:
            raise RBError(f"vwrays: bad view option '{a}'")
        v.type = a[3]
        return 0
    if len(a) != 3:
        raise RBError(f"vwrays: bad view option '{a}'")
    if c == "p":
        v.vp = np.array([f(1), f(2), f(3)]); return 3
    if c == "d":
        v.vdir = np.array([f(1), f(2), f(3)]); v.vdist = 1.; return 3
    if c == "u":
        v.vup = np.array([f(1), f(2), f(3)]); return 3
    if c in "hvoasl":
        setattr(v, {"h": "horiz", "v": "vert", "o": "vfore", "a": "vaft", "s": "hoff", "l": "voff"}[c], f(1))
        return 1
    raise RBError(f"vwrays: bad view option '{a}'")


def view_from_file(path, v: View) -> bool:
    """viewfile() / sscanview(): lines `VIEW= ...` or `rvu|rpict|... -v...` carry view options."""
    found = False
    for line in open(path, "r", errors="replace"):
        m = re.match(r"^\s*(VIEW=|rvu\b|rpict\b|rview\b|rtpict\b|vwright\b|pinterp\b)(.*)$", line)
        if not m:
            continue
        words = m.group(2).split()
        k = 0
        while k < len(words):
            if words[k].startswith("-v") and len(words[k]) >= 3 and words[k][2] in "tpduhvoasl":
                k += 1 + get_view_opts(v, words[k:])
                found = True
            else:
                k += 1
    return found


def set_view(v: View):
    """setview(): derive hvec / vvec / hn2 / vn2; raises on an illegal view"""
    if v.vfore < -FTINY or v.vaft < -FTINY or (v.vaft > FTINY and v.vaft <= v.vfore):
        raise RBError("vwrays: illegal fore/aft clipping plane")
    if v.vdist <= FTINY:
        raise RBError("vwrays: illegal view distance")
    ln, v.vdir = _normalize(v.vdir)
    v.vdist *= ln
    if v.vdist == 0.0:
        raise RBError("vwrays: zero view direction")
    ln, v.vup = _normalize(v.vup)
    if ln == 0.0:
        raise RBError("vwrays: zero view up vector")
    ln, v.hvec = _normalize(np.cross(v.vdir, v.vup))
    if ln == 0.0:
        raise RBError("vwrays: view up parallel to view direction")
    v.vvec = np.cross(v.hvec, v.vdir)
    if v.horiz <= FTINY:
        raise RBError("vwrays: illegal horizontal view size")
    if v.vert <= FTINY:
        raise RBError("vwrays: illegal vertical view size")
    t, PI = v.type, math.pi
    bad_h, bad_v = RBError("vwrays: illegal horizontal view size"), RBError("vwrays: illegal vertical view size")
    if t == "l":
        v.hn2, v.vn2 = v.horiz, v.vert
    elif t == "v":
        if v.horiz >= 180.0 - FTINY: raise bad_h
        if v.vert >= 180.0 - FTINY: raise bad_v
        v.hn2, v.vn2 = 2.0 * math.tan(v.horiz * (PI / 360.)), 2.0 * math.tan(v.vert * (PI / 360.))
    elif t == "c":
        if v.horiz > 360.0 + FTINY: raise bad_h
        if v.vert >= 180.0 - FTINY: raise bad_v
        v.hn2, v.vn2 = v.horiz * (PI / 180.0), 2.0 * math.tan(v.vert * (PI / 360.))
    elif t == "a":
        if v.horiz > 360.0 + FTINY: raise bad_h
        if v.vert > 360.0 + FTINY: raise bad_v
        v.hn2, v.vn2 = v.horiz * (PI / 180.0), v.vert * (PI / 180.0)
    elif t == "h":
        if v.horiz > 180.0 + FTINY: raise bad_h
        if v.vert > 180.0 + FTINY: raise bad_v
        v.hn2, v.vn2 = 2.0 * math.sin(v.horiz * (PI / 360.)), 2.0 * math.sin(v.vert * (PI / 360.))
    elif t == "s":
        if v.horiz >= 360.0 - FTINY: raise bad_h
        if v.vert >= 360.0 - FTINY: raise bad_v
        v.hn2 = 2. * math.sin(v.horiz * (PI / 360.)) / (1.0 + math.cos(v.horiz * (PI / 360.)))
        v.vn2 = 2. * math.sin(v.vert * (PI / 360.)) / (1.0 + math.cos(v.vert * (PI / 360.)))
    else:
        raise RBError("vwrays: unknown view type")
    if t not in "as":
        if t != "c":
            v.hvec = v.hvec * v.hn2
        v.vvec = v.vvec * v.vn2
    v.hn2 *= v.hn2
    v.vn2 *= v.vn2


def view_rays(v: View, x: np.ndarray, y: np.ndarray):
    """viewray() for arrays of image locations: returns (origins, directions, d) with d < 0 = no ray."""
    x = x + (v.hoff - 0.5)
    y = y + (v.voff - 0.5)
    n = x.shape[0]
    PI = math.pi
    vd, hv, vv = v.vdir, v.hvec, v.vvec
    aft = v.vaft - v.vfore if v.vaft > FTINY else 0.0

    def combo(z, xx, yy):
        return np.stack([(z * vd[k] + xx * hv[k]) + yy * vv[k] for k in range(3)], axis=1)

    if v.type == "l":
        org = np.stack([((v.vp[k] + v.vfore * vd[k]) + x * hv[k]) + y * vv[k] for k in range(3)], axis=1)
        return org, np.tile(vd, (n, 1)), np.full(n, aft)
    if v.type == "v":
        direc = np.stack([(vd[k] + x * hv[k]) + y * vv[k] for k in range(3)], axis=1)
        org = v.vp + direc * v.vfore
        d, direc = _normalize_rows(direc)
        return org, direc, (aft * d if v.vaft > FTINY else np.zeros(n))
    if v.type == "h":
        z = 1.0 - x * x * v.hn2 - y * y * v.vn2
        ok = z >= 0.0
        z = np.sqrt(np.where(ok, z, 0.0))
        direc = combo(z, x, y)
        return v.vp + direc * v.vfore, direc, np.where(ok, aft, -1.0)
    if v.type == "c":
        d = x * v.horiz * (PI / 180.0)
        direc = combo(np.cos(d), np.sin(d), y)
        org = v.vp + direc * v.vfore
        ln, direc = _normalize_rows(direc)
        return org, direc, (aft * ln if v.vaft > FTINY else np.zeros(n))
    if v.type == "a":
        x = x * ((1.0 / 180.0) * v.horiz)
        y = y * ((1.0 / 180.0) * v.vert)
        d = x * x + y * y
        ok = d <= 1.0
        d = np.sqrt(d)
        z = np.cos(PI * d)
        with np.errstate(divide="ignore", invalid="ignore"):
            s = np.where(d <= FTINY, PI, np.sqrt(1.0 - z * z) / np.where(d <= FTINY, 1.0, d))
        direc = combo(z, x * s, y * s)
        return v.vp + direc * v.vfore, direc, np.where(ok, aft, -1.0)
    if v.type == "s":
        x = x * math.sqrt(v.hn2)
        y = y * math.sqrt(v.vn2)
        d = x * x + y * y
        z = (1. - d) / (1. + d)
        direc = combo(z, x * (1. + z), y * (1. + z))
        return v.vp + direc * v.vfore, direc, np.full(n, aft)
    raise RBError("vwrays: unknown view type")


def vwrays_main(argv: Sequence[str], stdin: bytes | None = None, seed=None) -> bytes:
    """The vwrays command (argv[0] = program name)."""
    argv = [str(a) for a in argv]
    v = View()
    xr = yr = 512
    pa, pj, pd = 1.0, 0.0, 0.0
    outform, getdim, fromstdin, repeat = "a", False, False, 1
    i = 1
    usage = RBError("Usage: vwrays [ -i -u -f{a|f|d} -c rept | -d ] { view opts .. | picture [zbuf] }")
    while i < len(argv) and argv[i].startswith("-"):
        a = argv[i]
        c = a[1:2]
        if c == "f":
            if a[2:3] not in ("a", "f", "d") or a[2:3] == "":
                raise usage
            outform = a[2]
        elif c == "v":
            if a[2:3] == "f":
                if not view_from_file(argv[i + 1], v):
                    raise RBError(f"{argv[i + 1]}: no view in file")
                i += 1
            else:
                i += get_view_opts(v, argv[i:])
        elif c == "d":
            getdim = True
        elif c == "x":
            xr = int(argv[i + 1]); i += 1
            if xr <= 0: raise RBError("vwrays: bad x resolution")
        elif c == "y":
            yr = int(argv[i + 1]); i += 1
            if yr <= 0: raise RBError("vwrays: bad y resolution")
        elif c == "c":
            repeat = max(1, int(argv[i + 1])); i += 1
        elif c == "p":
            val = float(argv[i + 1]); i += 1
            if a[2:3] == "a": pa = val
            elif a[2:3] == "j": pj = val
            elif a[2:3] == "d": pd = val
            else: raise usage
        elif c == "i":
            fromstdin = True
        elif c == "u":
            pass
        else:
            raise usage
        i += 1
    if i < len(argv):
        raise RBError("vwrays: a picture / depth buffer as the view source is not built (give view options)")
    if pd > FTINY:
        raise RBError("vwrays: the depth-of-field aperture (-pd) is not built")
    set_view(v)
    va = math.sqrt(v.vn2 / v.hn2)                      # viewaspect()
    if pa <= FTINY:                                    # normaspect()
        pa = va * xr / yr
    elif va * xr > pa * yr:
        xr = int(yr / va * pa + .5)
    else:
        yr = int(xr * va / pa + .5)
    if getdim:
        return (f"-x {xr} -y {yr}" + (" -ld+" if v.vaft > FTINY else "") + "\n").encode()
    import os
    from . import _lib
    if not fromstdin and not os.environ.get("RB_VWRAYS_HOST") and _lib.device_count() > 0:
        # the rays of a whole view come from the device kernel (C ABI rb_view_rays: one thread per ray, same
        # expressions); the numpy code below stays for pixel lists on stdin and for machines without a GPU
        ctx = _lib.Context(0)
        try:
            rays = ctx.view_rays(v, xr, yr, repeat, pj, 0 if seed is None else int(seed))
        finally:
            ctx.close()
        return _format_rays(rays, outform)
    if fromstdin:
        vals = np.array((stdin or b"").split(), dtype=np.float64)
        vals = vals[:(vals.size // 2) * 2].reshape(-1, 2)
        lx = (vals[:, 0] + .5) / xr
        ly = (vals[:, 1] + .5) / yr
    else:
        sx, sy = np.meshgrid(np.arange(xr), np.arange(yr))        # scanlines from the top: -Y yr +X xr
        lx = (sx.ravel() + .5) / xr
        ly = ((yr - 1 - sy.ravel()) + .5) / yr
    if repeat > 1:
        lx, ly = np.repeat(lx, repeat), np.repeat(ly, repeat)
    if pj > FTINY:
        rng = np.random.default_rng(seed)
        lx = lx + pj * (.5 - rng.random(lx.shape[0])) / xr
        ly = ly + pj * (.5 - rng.random(ly.shape[0])) / yr
    org, direc, d = view_rays(v, lx, ly)
    bad = d < -FTINY
    scale = np.where(d > FTINY, d, 1.0)
    direc = direc * scale[:, None]
    org = np.where(bad[:, None], 0.0, org)
    direc = np.where(bad[:, None], 0.0, direc)
    rays = np.concatenate([org, direc], axis=1)
    return _format_rays(rays, outform)


def _format_rays(rays, outform) -> bytes:
    if outform == "d":
        return np.ascontiguousarray(rays, dtype=np.float64).tobytes()
    if outform == "f":
        return np.ascontiguousarray(rays, dtype=np.float32).tobytes()
    return "".join("%.5e %.5e %.5e %.5e %.5e %.5e\n" % tuple(r) for r in rays).encode()


def view_from_args(view_args: Sequence[str], xres: int, yres: int, pixel_aspect: float = 1.0):
    """View options -> (View after set_view(), xres, yres after the aspect normalisation of vwrays): what
    Context.view_rays() takes to generate the rays of a view straight into device memory."""
    v = View()
    words = [str(a) for a in view_args]
    k = 0
    while k < len(words):
        if words[k] == "-vf":
            if not view_from_file(words[k + 1], v):
                raise RBError(f"{words[k + 1]}: no view in file")
            k += 2
        else:
            k += 1 + get_view_opts(v, words[k:])
    set_view(v)
    va = math.sqrt(v.vn2 / v.hn2)
    pa = pixel_aspect
    if pa <= FTINY:
        pass
    elif va * xres > pa * yres:
        xres = int(yres / va * pa + .5)
    else:
        yres = int(xres * va / pa + .5)
    return v, xres, yres


def vwrays(pixpos: bytes | None = None, unbuf: bool = False, outform: str = "a", ray_count: int = 1,
           pixel_jitter: float = 0, pixel_diameter: float = 0, pixel_aspect: float = 1, xres: int = 512, yres: int = 512,
           dimensions: bool = False, view: Sequence[str] | None = None, pic=None, zbuf=None) -> bytes:
    """Same call as pyradiance.vwrays (src/pyradiance/util.py:1025-1066)."""
    cmd = ["vwrays"]
    if pixpos is not None:
        cmd.append("-i")
    if unbuf:
        cmd.append("-u")
    if outform != "a":
        cmd.append(f"-f{outform}")
    cmd += ["-c", str(ray_count), "-pj", str(pixel_jitter), "-pd", str(pixel_diameter), "-pa", str(pixel_aspect),
            "-x", str(xres), "-y", str(yres)]
    if dimensions:
        cmd.append("-d")
    if view is not None:
        cmd.extend(str(a) for a in view)
    elif pic is not None:
        cmd.append(str(pic))
        if zbuf is not None:
            cmd.append(str(zbuf))
    else:
        raise ValueError("Either view or pic should be provided.")
    return vwrays_main(cmd, pixpos)
