"""Seeded synthetic scenes and sensor sets for the parity tests and bench.py.

These are the inputs SURVEY.md 8(d) names (S-office-100k, S-building-1M and
small variants): a Radiance text scene (.rad) is written, then frozen into an
octree either by our own builder (``rb_oconv``, the product path used by
bench.py) or by the reference ``oconv`` (tests only).  Everything is
deterministic in the seed.
"""
from __future__ import annotations

import io
import os
from pathlib import Path

import numpy as np

MATERIALS = """\
void plastic floor_mat
0
0
5 .2 .2 .2 0 0

void plastic wall_mat
0
0
5 .5 .5 .5 0 0

void plastic ceil_mat
0
0
5 .8 .8 .8 0 0

void plastic furn_a
0
0
5 .3 .3 .3 0 0

void plastic furn_b
0
0
5 .45 .45 .45 0 0

void plastic furn_c
0
0
5 .6 .55 .5 0 0

void metal trim_mat
0
0
5 .6 .6 .6 .8 0

void glass win_glass
0
0
3 .654 .654 .654

"""

SKY = """\
void glow skyglow
0
0
4 1 1 1 0

skyglow source sky
0
0
4 0 0 1 180

void glow groundglow
0
0
4 1 1 1 0

groundglow source ground
0
0
4 0 0 -1 180

"""

SUN = """\
void light solar
0
0
3 6e6 6e6 6e6

solar source sun
0
0
4 0.2 -0.6 0.77 0.533

"""


def _poly(out, mod, name, verts):
    out.write(f"{mod} polygon {name}\n0\n0\n{3 * len(verts)}\n")
    for v in verts:
        out.write(f" {v[0]:.9g} {v[1]:.9g} {v[2]:.9g}\n")
    out.write("\n")


def _box(out, mod, name, lo, hi):
    x0, y0, z0 = lo
    x1, y1, z1 = hi
    # six outward-facing quads (counter-clockwise seen from outside)
    _poly(out, mod, name + ".b", [(x0, y0, z0), (x0, y1, z0), (x1, y1, z0), (x1, y0, z0)])
    _poly(out, mod, name + ".t", [(x0, y0, z1), (x1, y0, z1), (x1, y1, z1), (x0, y1, z1)])
    _poly(out, mod, name + ".s", [(x0, y0, z0), (x1, y0, z0), (x1, y0, z1), (x0, y0, z1)])
    _poly(out, mod, name + ".n", [(x0, y1, z0), (x0, y1, z1), (x1, y1, z1), (x1, y1, z0)])
    _poly(out, mod, name + ".w", [(x0, y0, z0), (x0, y0, z1), (x0, y1, z1), (x0, y1, z0)])
    _poly(out, mod, name + ".e", [(x1, y0, z0), (x1, y1, z0), (x1, y1, z1), (x1, y0, z1)])


def office_floor(out, rng, z0, nboxes, tag="f0", W=40.0, D=25.0, H=3.0, nwin=8,
                 frac_sphere=0.02, frac_cyl=0.02):
    """One office floor: shell with `nwin` south windows + clutter.  Returns
    the number of surfaces written."""
    n = 0
    z1 = z0 + H
    _poly(out, "floor_mat", f"{tag}.floor", [(0, 0, z0), (W, 0, z0), (W, D, z0), (0, D, z0)]); n += 1
    _poly(out, "ceil_mat", f"{tag}.ceil", [(0, 0, z1), (0, D, z1), (W, D, z1), (W, 0, z1)]); n += 1
    _poly(out, "wall_mat", f"{tag}.north", [(0, D, z0), (W, D, z0), (W, D, z1), (0, D, z1)]); n += 1
    _poly(out, "wall_mat", f"{tag}.west", [(0, 0, z0), (0, D, z0), (0, D, z1), (0, 0, z1)]); n += 1
    _poly(out, "wall_mat", f"{tag}.east", [(W, 0, z0), (W, 0, z1), (W, D, z1), (W, D, z0)]); n += 1
    # south facade (y = 0) with window openings
    sill, head = z0 + 0.9, z0 + 2.7
    bay = W / nwin
    _poly(out, "wall_mat", f"{tag}.south.sill", [(0, 0, z0), (0, 0, sill), (W, 0, sill), (W, 0, z0)]); n += 1
    _poly(out, "wall_mat", f"{tag}.south.head", [(0, 0, head), (0, 0, z1), (W, 0, z1), (W, 0, head)]); n += 1
    xs = [0.0]
    for i in range(nwin):
        a, b = i * bay + 0.15 * bay, i * bay + 0.85 * bay
        _poly(out, "win_glass", f"{tag}.win{i}", [(a, 0, sill), (a, 0, head), (b, 0, head), (b, 0, sill)]); n += 1
        xs += [a, b]
    xs.append(W)
    for i in range(0, len(xs), 2):
        a, b = xs[i], xs[i + 1]
        _poly(out, "wall_mat", f"{tag}.south.pier{i // 2}", [(a, 0, sill), (a, 0, head), (b, 0, head), (b, 0, sill)]); n += 1
    mats = ["furn_a", "furn_b", "furn_c", "trim_mat"]
    nsph = int(round(nboxes * 6 * frac_sphere))
    ncyl = int(round(nboxes * 6 * frac_cyl))
    nb = max(0, nboxes - (nsph + ncyl) // 6)
    # two layers: furniture below the work plane, services above it
    cx = rng.uniform(0.4, W - 0.4, nb)
    cy = rng.uniform(0.4, D - 0.4, nb)
    sx = rng.uniform(0.03, 0.25, nb)
    sy = rng.uniform(0.03, 0.25, nb)
    low = rng.random(nb) < 0.6
    zb = np.where(low, z0 + 0.01 + rng.uniform(0, 0.45, nb), z0 + 0.9 + rng.uniform(0, 1.6, nb))
    hh = np.where(low, rng.uniform(0.03, 0.28, nb), rng.uniform(0.03, 0.4, nb))
    mi = rng.integers(0, len(mats), nb)
    for i in range(nb):
        _box(out, mats[mi[i]], f"{tag}.bx{i}", (cx[i] - sx[i], cy[i] - sy[i], zb[i]),
             (cx[i] + sx[i], cy[i] + sy[i], zb[i] + hh[i]))
    n += 6 * nb
    for i in range(nsph):
        c = (rng.uniform(0.5, W - 0.5), rng.uniform(0.5, D - 0.5), z0 + rng.uniform(1.0, 2.6))
        r = rng.uniform(0.04, 0.2)
        out.write(f"{mats[i % 4]} sphere {tag}.sp{i}\n0\n0\n4 {c[0]:.9g} {c[1]:.9g} {c[2]:.9g} {r:.9g}\n\n")
    n += nsph
    for i in range(ncyl):
        p = np.array([rng.uniform(0.5, W - 0.5), rng.uniform(0.5, D - 0.5), z0 + rng.uniform(1.0, 2.5)])
        d = rng.normal(size=3)
        d /= np.linalg.norm(d)
        q = p + d * rng.uniform(0.2, 0.45)
        q[2] = min(max(q[2], z0 + 0.9), z1 - 0.05)
        r = rng.uniform(0.01, 0.06)
        out.write(f"{mats[i % 4]} cylinder {tag}.cy{i}\n0\n0\n7 {p[0]:.9g} {p[1]:.9g} {p[2]:.9g} "
                  f"{q[0]:.9g} {q[1]:.9g} {q[2]:.9g} {r:.9g}\n\n")
    n += ncyl
    return n



_TNAZ = (30, 30, 24, 24, 18, 12, 6)


def reinhart_suns(mf=1):
    """Unit vectors to the centres of the Reinhart MF:n sky patches in the bin
    order of reinhart.cal's rbin (sky bins 1..144*mf*mf+1; bin 0 is the ground):
    the sun positions of a 5-phase direct-sun matrix (BASELINE config 5)."""
    alpha = 90.0 / (7 * mf + 0.5)
    dirs = []
    for row in range(7 * mf):
        n = mf * _TNAZ[row // mf]
        alt = np.radians((row + 0.5) * alpha)
        for k in range(n):
            azi = np.radians(k * 360.0 / n)
            dirs.append((np.sin(azi) * np.cos(alt), np.cos(azi) * np.cos(alt), np.sin(alt)))
    dirs.append((0.0, 0.0, 1.0))
    return np.array(dirs)


def write_suns(out, mf=1, modifier="solar", radiance=1e6, angle=0.533):
    """`light` material + one `source` per Reinhart patch centre, all sharing one
    modifier (the form rcontrib's sun-coefficient runs use)."""
    out.write(f"void light {modifier}\n0\n0\n3 {radiance:g} {radiance:g} {radiance:g}\n\n")
    for i, d in enumerate(reinhart_suns(mf)):
        out.write(f"{modifier} source sun{i}\n0\n0\n4 {d[0]:.8f} {d[1]:.8f} {d[2]:.8f} {angle:g}\n\n")

def write_office(path, npolys=100_000, floors=1, seed=1234, sun=False, curved=True):
    """Write the S-office (floors=1) / S-building (floors=10) scene; returns
    the surface count."""
    rng = np.random.default_rng(seed)
    out = io.StringIO()
    out.write(f"# synthetic office: {npolys} surfaces, {floors} floor(s), seed {seed}\n")
    out.write(MATERIALS)
    out.write(SKY)
    if sun:
        out.write(SUN)
    per_floor = npolys // floors
    shell = 7 + 8 + 9
    nboxes = max(0, (per_floor - shell) // 6)
    n = 0
    for f in range(floors):
        n += office_floor(out, rng, 3.3 * f, nboxes, tag=f"f{f}", **({} if curved else {"frac_sphere": 0.0, "frac_cyl": 0.0}))
    Path(path).write_text(out.getvalue())
    return n


def office_sensors(nsensors, floors=1, seed=42, W=40.0, D=25.0, height=0.8):
    """Jittered grid of upward-facing sensors on the work plane of each floor:
    float64 [n, 6] rows origin, direction."""
    rng = np.random.default_rng(seed)
    per = nsensors // floors
    rows = []
    for f in range(floors):
        cnt = per if f < floors - 1 else nsensors - per * (floors - 1)
        nx = int(np.ceil(np.sqrt(cnt * W / D)))
        ny = int(np.ceil(cnt / nx))
        gx, gy = np.meshgrid(np.arange(nx), np.arange(ny), indexing="xy")
        gx = gx.ravel()[:cnt]
        gy = gy.ravel()[:cnt]
        x = 0.5 + (gx + rng.random(cnt)) * (W - 1.0) / nx
        y = 0.5 + (gy + rng.random(cnt)) * (D - 1.0) / ny
        z = np.full(cnt, 3.3 * f + height)
        rows.append(np.stack([x, y, z, np.zeros(cnt), np.zeros(cnt), np.ones(cnt)], axis=1))
    return np.ascontiguousarray(np.concatenate(rows, axis=0), dtype=np.float64)


def random_rays(n, seed=7, lo=(0.3, 0.3, 0.1), hi=(39.7, 24.7, 2.9)):
    """Uniform random origins in a box with uniform random directions."""
    rng = np.random.default_rng(seed)
    o = rng.uniform(lo, hi, size=(n, 3))
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    return np.ascontiguousarray(np.concatenate([o, d], axis=1), dtype=np.float64)


def build_octree(rad_path, oct_path, use_reference_oconv=None):
    """Freeze a text scene into an octree.  Product path: our own builder.
    Tests may pass the path of the reference oconv binary instead."""
    if use_reference_oconv:
        import subprocess
        with open(oct_path, "wb") as f:
            subprocess.run([use_reference_oconv, "-f", os.path.basename(os.fspath(rad_path))], check=True, stdout=f,
                           cwd=os.path.dirname(os.path.abspath(os.fspath(rad_path))))      # nested .oct / .rtm are relative
        return
    from ._lib import oconv_file
    oconv_file(rad_path, oct_path)
