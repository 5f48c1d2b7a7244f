"""Golden fixtures that pin the surfaces and tie rules no other fixture exercises
(VERDICT round 1 "pin the untested geometry").

TEST INFRASTRUCTURE.  Run in the build container (needs oracle/_ref):

    python tests/golden/make_golden_geom.py

Writes tests/golden/geom/{curved.rad,curved.oct,curved_text.oct,coinc.rad,coinc.oct,coinc_fine.oct}
and tests/golden/geom.npz.

curved.rad -- every member of the cone family (rt/o_cone.c:17-146, common/cone.c:44-153): cone
  (frustum and one with an apex), cup (frustum, seen from inside and outside), cylinder and tube
  (axis-parallel and oblique), ring (with and without a hole, tilted), sphere and bubble
  (rt/sphere.c:16-83; rays start inside and outside).  Rays: aimed at the bodies from outside, started
  inside the hollow ones, and aimed at the end-cap rims with offsets of 1e-3 ... 1e-9 across the rim
  (the `0 <= b <= al` and radius tests).  curved_text.oct is the same scene NOT frozen
  (common/readoct.c:90-100: the loader must read curved.rad through the octree's file list).
coinc.rad -- coincident and nearly coincident surfaces for rayreject() (rt/raytrace.c:535-575):
  material vs void, opaque vs transparent, front vs back, modifier order, identical modifiers,
  offsets below and above FTINY, a three-step chain, partly overlapping quads, a ring on a quad;
  each case in both definition orders.  coinc_fine.oct is the same scene under `oconv -n 1 -r 256`
  so that the surfaces straddle many leaves (re-tests across leaves).
Known answers: the UNMODIFIED reference rtrace (-ab 0 -dt 0 -dj 0 -dc 1 -av .1 .1 .1): surface and
modifier names (ascii run), distance, normal and value as float64 (binary run).
"""
import os
import subprocess
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent.parent))
from oracle import refrun  # noqa: E402

S = HERE / "geom"
S.mkdir(exist_ok=True)
env = dict(os.environ, RAYPATH=f".:{refrun.LIB}")

MATS = """void plastic grey
0
0
5 .5 .5 .5 0 0

void plastic red
0
0
5 .7 .2 .2 .03 .08

void plastic green
0
0
5 .2 .7 .2 0 0

void plastic blue
0
0
5 .2 .2 .7 0 0

void metal steel
0
0
5 .6 .6 .65 .8 .1

void glass pane
0
0
3 .8 .85 .8

void light sunl
0
0
3 5000 5000 4500

sunl source sun
0
0
4 .3 -.2 .93 1.5

void glow skyg
0
0
4 .8 .9 1.2 0

skyg source sky
0
0
4 0 0 1 180

skyg source ground
0
0
4 0 0 -1 180

"""

CURVED = MATS + """grey polygon floor
0
0
12 -4 -4 -.5  14 -4 -.5  14 12 -.5  -4 12 -.5

red cone frustum
0
0
8 0 0 0  0 0 2  1 .3

green cone spike
0
0
8 4 0 0  4.3 .2 2.2  .9 0

blue cup bowl
0
0
8 8 0 0  8 0 1.5  .4 1.2

steel cup funnel
0
0
8 11 0 1.8  11.4 .3 0  1.1 .15

red cylinder post
0
0
7 0 4 0  0 4 2  .7

green cylinder strut
0
0
7 2 4 0  3.5 5 1.8  .4

blue tube pipe
0
0
7 8 4 0  8 4 2.5  1

steel tube duct
0
0
7 10.5 3.5 .2  12.5 5 1.6  .6

red ring washer
0
0
8 4 4 1  .3 .2 1  .3 1.1

green ring disc
0
0
8 5.5 6.5 .7  -.2 .5 .8  0 .8

pane ring porthole
0
0
8 4 4 1.6  .3 .2 1  0 .9

blue sphere ball
0
0
4 0 8 1 1

steel bubble cave
0
0
4 4 8.5 1 1.2

red sphere pearl
0
0
4 4.2 8.4 .9 .35

green bubble void_ball
0
0
4 8 8 1.2 .8

pane sphere marble
0
0
4 11 8 .8 .7

"""

COINC_MATS = MATS + """void plastic early
0
0
5 .4 .4 .1 0 0

void plastic late
0
0
5 .1 .4 .4 0 0

"""


def quad(mod, name, x0, y0, x1, y1, z, flip=False):
    v = [(x0, y0, z), (x1, y0, z), (x1, y1, z), (x0, y1, z)]
    if flip:
        v = v[::-1]
    return f"{mod} polygon {name}\n0\n0\n12 " + "  ".join(" ".join(f"{c:.10g}" for c in p) for p in v) + "\n\n"


def coinc_scene():
    s = COINC_MATS
    t = 0          # tile counter: tile i covers x in [2i, 2i+1], y in [0, 1]

    def tile(parts):
        nonlocal s, t
        for k, (mod, dz, flip, ext) in enumerate(parts):
            x0, x1, y0, y1 = 2 * t + ext[0], 2 * t + 1 + ext[1], ext[2], 1 + ext[3]
            s += quad(mod, f"t{t}_{k}_{mod}", x0, y0, x1, y1, dz, flip)
        t += 1

    E = (0, 0, 0, 0)
    tile([("grey", 0, False, E), ("void", 0, False, E)])            # material beats none ...
    tile([("void", 0, False, E), ("grey", 0, False, E)])            # ... in either order
    tile([("pane", 0, False, E), ("red", 0, False, E)])             # opaque beats transparent
    tile([("red", 0, False, E), ("pane", 0, False, E)])
    tile([("green", 0, False, E), ("green", 0, True, E)])           # front beats back (seen from both sides)
    tile([("green", 0, True, E), ("green", 0, False, E)])
    tile([("early", 0, False, E), ("late", 0, False, E)])           # later modifier definition wins
    tile([("late", 0, False, E), ("early", 0, False, E)])
    tile([("blue", 0, False, E), ("blue", 0, False, E)])            # identical modifier: first tested stays
    tile([("early", 0, False, E), ("late", 5e-7, False, E)])        # offset below FTINY: still a tie
    tile([("late", 0, False, E), ("early", 5e-7, False, E)])
    tile([("early", 0, False, E), ("late", 1.5e-6, False, E)])      # offset above FTINY: nearer one wins
    tile([("late", 0, False, E), ("early", -1.5e-6, False, E)])
    tile([("late", 0, False, E), ("red", 8e-7, False, E), ("early", 1.6e-6, False, E)])   # chain of ties
    tile([("early", 0, False, E), ("red", 8e-7, False, E), ("late", 1.6e-6, False, E)])
    tile([("early", 0, False, (0, -.4, 0, 0)), ("late", 0, False, (.3, 0, 0, -.2)), ("pane", 0, False, (0, 0, .5, 0))])  # partial overlaps
    tile([("void", 0, False, E), ("pane", 0, False, E), ("void", 0, True, E)])
    # a ring and a sphere's pole touching a quad
    s += quad("grey", "under_ring", 2 * t, 0, 2 * t + 1, 1, 0)
    s += f"late ring on_quad\n0\n0\n8 {2 * t + .5} .5 0  0 0 1  .1 .45\n\n"
    t += 1
    s += quad("grey", "under_ball", 2 * t, 0, 2 * t + 1, 1, 0)
    s += f"late sphere on_quad_ball\n0\n0\n4 {2 * t + .5} .5 .3 .3\n\n"
    t += 1
    return s, t


def ref_rtrace(octree, rays, cwd):
    det = ["-ab", "0", "-dt", "0", "-dj", "0", "-dc", "1", "-av", ".1", ".1", ".1", "-st", "1", "-lr", "6", "-lw", "1e-3"]
    a = subprocess.run([str(refrun.BIN / "rtrace"), "-h", "-fda", "-osm"] + det + [octree], cwd=cwd, env=env,
                       capture_output=True, input=rays.tobytes())
    assert a.returncode == 0, a.stderr.decode()
    rows = [ln.split("\t") for ln in a.stdout.decode().splitlines()]
    b = subprocess.run([str(refrun.BIN / "rtrace"), "-h", "-fdd", "-oLNv"] + det + [octree], cwd=cwd, env=env,
                       capture_output=True, input=rays.tobytes())
    assert b.returncode == 0, b.stderr.decode()
    d = np.frombuffer(b.stdout, dtype=np.float64).reshape(len(rays), 7)
    return {"surf": np.array([q[0] for q in rows]), "mod": np.array([q[1] for q in rows]), "dist": d[:, 0].copy(),
            "norm": d[:, 1:4].copy(), "value": d[:, 4:7].copy()}, np.array(det)


def unit(v):
    v = np.asarray(v, dtype=float)
    return v / np.linalg.norm(v, axis=-1, keepdims=True)


def curved_rays(rng):
    bodies = {  # name: (centre, half extent for aiming)
        "frustum": ((0, 0, 1), (1.1, 1.1, 1.1)), "spike": ((4.15, .1, 1.1), (1, 1, 1.2)), "bowl": ((8, 0, .75), (1.3, 1.3, .9)),
        "funnel": ((11.2, .15, .9), (1.2, 1.2, 1)), "post": ((0, 4, 1), (.8, .8, 1.1)), "strut": ((2.75, 4.5, .9), (1, .8, 1)),
        "pipe": ((8, 4, 1.25), (1.1, 1.1, 1.4)), "duct": ((11.5, 4.25, .9), (1.2, 1, .9)), "washer": ((4, 4, 1.3), (1.2, 1.2, .4)),
        "disc": ((5.5, 6.5, .7), (.9, .9, .5)), "ball": ((0, 8, 1), (1.1, 1.1, 1.1)), "cave": ((4, 8.5, 1), (1.3, 1.3, 1.3)),
        "void_ball": ((8, 8, 1.2), (.9, .9, .9)), "marble": ((11, 8, .8), (.8, .8, .8)),
    }
    rays = []
    for c, h in bodies.values():                      # from outside
        n = 260
        tgt = np.array(c) + rng.uniform(-1, 1, (n, 3)) * np.array(h)
        org = tgt + unit(rng.normal(size=(n, 3))) * rng.uniform(1.5, 6, (n, 1))
        org[:, 2] = np.abs(org[:, 2]) + .05
        rays.append(np.concatenate([org, unit(tgt - org)], 1))
    for name in ("bowl", "funnel", "pipe", "duct", "cave", "void_ball", "ball", "marble", "frustum", "post"):   # from inside
        c, h = bodies[name]
        n = 200
        org = np.array(c) + rng.uniform(-.3, .3, (n, 3)) * np.array(h)
        rays.append(np.concatenate([org, unit(rng.normal(size=(n, 3)))], 1))
    # end-cap rims: points on the rim circles of the axis-parallel bodies, shifted across the rim
    rims = [((0, 0, 0), 1.0), ((0, 0, 2), .3), ((8, 0, 0), .4), ((8, 0, 1.5), 1.2), ((0, 4, 0), .7), ((0, 4, 2), .7),
            ((8, 4, 0), 1.0), ((8, 4, 2.5), 1.0)]
    for (cx, cy, cz), r in rims:
        for eps in (1e-3, 1e-5, 1e-6, 1e-7, 1e-9, 0.0, -1e-9, -1e-7, -1e-6, -1e-5, -1e-3):
            for k in range(6):
                a = rng.uniform(0, 2 * np.pi)
                # a point on the side surface, eps above / below the cap plane
                p = np.array([cx + r * np.cos(a), cy + r * np.sin(a), cz + eps])
                org = p + np.array([np.cos(a + .3) * 3, np.sin(a + .3) * 3, rng.uniform(-1.5, 1.5)])
                org[2] = max(org[2], -.3)
                rays.append(np.concatenate([org, unit(p - org)])[None])
    # ring radii: across the inner and outer edge of the tilted washer
    n = unit([.3, .2, 1]); u = unit(np.cross(n, [1, 0, 0])); v = np.cross(n, u)
    for rr in (.3, 1.1):
        for eps in (1e-3, 1e-6, 1e-8, 0.0, -1e-8, -1e-6, -1e-3):
            for k in range(6):
                a = rng.uniform(0, 2 * np.pi)
                p = np.array([4, 4, 1]) + (rr + eps) * (np.cos(a) * u + np.sin(a) * v)
                org = p + unit(n * rng.choice([-1, 1]) + rng.normal(size=3) * .4) * 2.5
                rays.append(np.concatenate([org, unit(p - org)])[None])
    return np.concatenate(rays, 0)


def coinc_rays(rng, ntiles):
    rays = []
    for t in range(ntiles):
        n = 160
        tgt = np.stack([rng.uniform(2 * t - .15, 2 * t + 1.15, n), rng.uniform(-.15, 1.15, n), np.zeros(n)], 1)
        side = np.where(rng.uniform(size=n) < .6, 1.0, -1.0)
        d = unit(np.stack([rng.normal(size=n) * .5, rng.normal(size=n) * .5, -side * rng.uniform(.2, 1, n)], 1))
        org = tgt - d * rng.uniform(.5, 3, (n, 1))
        rays.append(np.concatenate([org, d], 1))
        # straight down / up through the middle and along the edges
        for x, y in ((.5, .5), (0, .5), (1, .5), (.5, 0), (.3, .8), (.6, .999999), (1e-7, .4)):
            for sz in (1.0, -1.0):
                rays.append(np.array([[2 * t + x, y, 2 * sz, 0, 0, -sz]]))
    return np.concatenate(rays, 0)


def main():
    rng = np.random.default_rng(23)
    out = {}
    (S / "curved.rad").write_text(CURVED)
    for args, name in ((["-f"], "curved.oct"), ([], "curved_text.oct")):
        r = subprocess.run([str(refrun.BIN / "oconv")] + args + ["curved.rad"], cwd=S, env=env, capture_output=True)
        assert r.returncode == 0, r.stderr.decode()
        (S / name).write_bytes(r.stdout)
    rays = curved_rays(rng)
    g, det = ref_rtrace("curved.oct", rays, S)
    g2, _ = ref_rtrace("curved_text.oct", rays, S)
    # (a frozen octree keeps ~31 bits of every real, so the two runs agree only to ~1e-9)
    print("frozen vs text octree: surfaces differ on", int((g["surf"] != g2["surf"]).sum()), "rays, max distance difference",
          float(np.abs(g["dist"] - g2["dist"])[g["surf"] == g2["surf"]].max()))
    out["args"] = det
    out["curved_rays"] = rays
    for k, v in g.items():
        out["curved_" + k] = v
    for k, v in g2.items():
        out["curvedtext_" + k] = v
    names, counts = np.unique(g["surf"], return_counts=True)
    print("curved:", len(rays), "rays;", dict(zip(names.tolist(), counts.tolist())))
    txt, ntiles = coinc_scene()
    (S / "coinc.rad").write_text(txt)
    for args, name in ((["-f"], "coinc.oct"), (["-f", "-n", "1", "-r", "256"], "coinc_fine.oct")):
        r = subprocess.run([str(refrun.BIN / "oconv")] + args + ["coinc.rad"], cwd=S, env=env, capture_output=True)
        assert r.returncode == 0, r.stderr.decode()
        (S / name).write_bytes(r.stdout)
    rays = coinc_rays(rng, ntiles)
    out["coinc_rays"] = rays
    for tag, octf in (("coinc", "coinc.oct"), ("coincfine", "coinc_fine.oct")):
        g, _ = ref_rtrace(octf, rays, S)
        for k, v in g.items():
            out[f"{tag}_" + k] = v
        names, counts = np.unique(g["surf"], return_counts=True)
        print(tag + ":", len(rays), "rays;", dict(zip(names.tolist(), counts.tolist())))
    same = np.array_equal(out["coinc_surf"], out["coincfine_surf"])
    print("coinc vs coinc_fine identical surfaces:", same, int((out["coinc_surf"] != out["coincfine_surf"]).sum()))
    np.savez_compressed(HERE / "geom.npz", **out)
    print("wrote", HERE / "geom.npz")


if __name__ == "__main__":
    main()
