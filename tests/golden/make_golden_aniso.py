"""Golden fixture for the anisotropic materials plastic2 / metal2 / trans2 (rt/aniso.c),
SURVEY 8f row f4.

TEST INFRASTRUCTURE.  Run in the build container (needs oracle/_ref, the unmodified
reference oconv / rtrace / rcontrib built by oracle/Makefile):

    python tests/golden/make_golden_aniso.py

Writes tests/golden/aniso/{aniso.rad,aniso.oct,anisoxf.rad,anisoxf.oct} and
tests/golden/aniso.npz.  Scene: seven 2 m panels at z = 1 over a grey floor --
plastic2 (u along x), metal2 (u diagonal), trans2 (diffuse + specular transmission),
a plastic2 whose orientation vector is parallel to the normal (getacoords() "punting"
branch), a metal2 that is nearly a mirror lobe, a trans2 without specular transmission,
an ordinary plastic for reference -- plus a metal2 sphere (not flat: no source-size term
in the highlight); a distant sun, a local rectangular lamp under the ceiling height and
a lamp BELOW the panels (so the trans2 panels are lit from behind, seen from above, and
the other way round), a glow sky.  anisoxf.rad holds the same materials with a function
transform (-rz 35 -rx 20) on the orientation vector; it is only run through the GPU path
(the CPU restatement takes untransformed constants).

Deterministic part: -ab 0 -dt 0 -dj 0 -dc 1 -st 1 -av .02 .03 .04 (every highlight is
below the sampling threshold, so the value is diraniso() over the sources plus the
constant ambient term with the specular colour folded in), view rays from above and from
below, and -I sensors.  Stochastic part: -st 0 (agaussamp() rays on), 160 view rays x 1500
repetitions, mean and standard error of the reference per ray.
"""
import os
import subprocess
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent.parent))
from oracle import refrun  # noqa: E402

S = HERE / "aniso"
S.mkdir(exist_ok=True)
env = dict(os.environ, RAYPATH=f".:{refrun.LIB}")

MATS = """void plastic2 p2x
{ns} 1 0 0 .{xf}
0
6 .5 .3 .2 .08 .05 .25

void metal2 m2d
{ns} 1 1 0 .{xf}
0
6 .7 .6 .3 .8 .3 .08

void trans2 t2
{ns} 0 1 0 .{xf}
0
8 .6 .7 .8 .06 .1 .2 .7 .6

void plastic2 p2punt
4 0 0 1 .
0
6 .4 .5 .4 .1 .1 .3

void metal2 m2sharp
{ns} 1 0 .2 .{xf}
0
6 .8 .8 .8 .9 .02 .03

void trans2 t2diff
{ns} 1 0 0 .{xf}
0
8 .7 .7 .6 .03 .2 .1 .5 0

void plastic plain
0
0
5 .5 .5 .5 .05 .1

void metal2 m2ball
{ns} 0 0 1 .{xf}
0
6 .6 .5 .4 .7 .15 .05
"""

GEOM = """void plastic grey
0
0
5 .3 .3 .3 0 0

grey polygon floor
0
0
12 -2 -2 0  16 -2 0  16 6 0  -2 6 0

{panels}
m2ball sphere ball
0
0
4 7 5 2 .8

void light sunl
0
0
3 9000 9000 8000

sunl source sun
0
0
4 .3 -.4 .85 1.5

void light lampl
0
0
3 60 55 40

lampl polygon lamp
0
0
12 5 1 4  8 1 4  8 3 4  5 3 4

void light lowl
0
0
3 40 40 60

lowl polygon lowlamp
0
0
12 3 0 .2  3 2 .2  11 2 .2  11 0 .2

void glow skyg
0
0
4 .8 .9 1.2 0

skyg source sky
0
0
4 0 0 1 180
"""

names = ["p2x", "m2d", "t2", "p2punt", "m2sharp", "t2diff", "plain"]
panels = ""
for i, m in enumerate(names):
    x0 = 2 * i
    panels += f"{m} polygon panel_{m}\n0\n0\n12 {x0} 0 1  {x0 + 1.9} 0 1  {x0 + 1.9} 2 1  {x0} 2 1\n\n"


def sh(cmd, out=None, stdin=None):
    r = subprocess.run(cmd, cwd=S, env=env, capture_output=True, input=stdin)
    assert r.returncode == 0, r.stderr.decode()
    if out:
        (S / out).write_bytes(r.stdout)
    return r.stdout


(S / "aniso.rad").write_text(MATS.format(ns=4, xf="") + GEOM.format(panels=panels))
(S / "anisoxf.rad").write_text(MATS.format(ns=8, xf=" -rz 35 -rx 20") + GEOM.format(panels=panels))
sh([str(refrun.BIN / "oconv"), "-f", "aniso.rad"], out="aniso.oct")
sh([str(refrun.BIN / "oconv"), "-f", "anisoxf.rad"], out="anisoxf.oct")

rng = np.random.default_rng(11)
n = 2400
# view rays aimed at points on the panel strip / the ball, from above (two thirds) and from below
tgt = np.stack([rng.uniform(-0.5, 14.5, n), rng.uniform(-0.3, 2.3, n), np.full(n, 1.0)], 1)
ball = rng.random(n) < 0.12
tgt[ball] = np.array([7, 5, 2]) + rng.normal(size=(ball.sum(), 3)) * 0.4
org = tgt + np.stack([rng.uniform(-3, 3, n), rng.uniform(-3, 3, n), rng.uniform(0.6, 3, n)], 1)
below = (rng.random(n) < 0.33) & ~ball
org[below, 2] = rng.uniform(0.25, 0.9, below.sum())
d = tgt - org
d /= np.linalg.norm(d, axis=1, keepdims=True)
rays = np.concatenate([org, d], 1)
det = ["-ab", "0", "-dt", "0", "-dj", "0", "-dc", "1", "-st", "1", "-av", ".02", ".03", ".04"]
out = {"rays": rays, "args": np.array(det)}


def trace(octf, rays, args, spec="vLsm"):
    r = sh([str(refrun.BIN / "rtrace"), "-h", "-fda", "-o" + spec] + args + [octf], stdin=rays.tobytes()).decode()
    rows = [ln.split("\t") for ln in r.splitlines()]
    return (np.array([[float(x) for x in q[0:3]] for q in rows]), np.array([float(q[3]) for q in rows]),
            np.array([q[4] for q in rows]), np.array([q[5] for q in rows]))


for tag, octf in (("", "aniso.oct"), ("xf_", "anisoxf.oct")):
    v, L, s, m = trace(octf, rays, det)
    out[tag + "value"], out[tag + "dist"], out[tag + "surf"], out[tag + "mod"] = v, L, s, m
    print(tag or "plain", {k: int((m == k).sum()) for k in names + ["m2ball", "grey"]})

# -I sensors just above and just below the trans2 panels see the lamps through them
sens = np.array([[x, y, z, 0, 0, dz] for x in (4.3, 5.2, 10.4, 11.6) for y in (0.4, 1.1, 1.7)
                 for z, dz in ((0.6, 1.0), (1.6, -1.0))], dtype=float)
out["sensors"] = sens
out["irrad"] = refrun.rtrace(S / "aniso.oct", sens, ["-I"] + det, outform="d").reshape(-1, 3)

# stochastic: highlights sampled (-st 0), no ambient bounce; mean over repetitions
nst, reps = 160, 1500
pick = np.flatnonzero(np.isin(out["mod"], ["p2x", "m2d", "t2", "m2sharp", "m2ball", "p2punt"]))[:nst]
st = ["-ab", "0", "-dt", "0", "-dj", "0", "-dc", "1", "-st", "0", "-av", ".02", ".03", ".04"]
big = np.tile(rays[pick], (reps, 1))
v = refrun.rtrace(S / "aniso.oct", big, st, outform="d").reshape(reps, len(pick), 3)
out["st_pick"] = pick
out["st_args"] = np.array(st)
out["st_mean"] = v.mean(0)
out["st_sem"] = v.std(0, ddof=1) / np.sqrt(reps)
print("stochastic: mean", out["st_mean"].mean(0), "relative sem (median)",
      np.median(out["st_sem"][:, 1] / np.maximum(out["st_mean"][:, 1], 1e-9)))

# rcontrib coefficients at -ab 1: the sky's share of what the panels reflect / transmit (tracked: skyg, lampl)
rc = ["-ab", "1", "-ad", "256", "-lw", "1e-3", "-dj", "0", "-st", "0", "-m", "skyg", "-m", "lampl", "-m", "lowl"]
rr = np.tile(rays[pick[:60]], (200, 1))
m = refrun.rcontrib(S / "aniso.oct", rr, rc).reshape(200, 60, 3, 3)
out["rc_args"] = np.array(rc[:10])
out["rc_mean"] = m.mean(0)
out["rc_sem"] = m.std(0, ddof=1) / np.sqrt(200)
np.savez_compressed(HERE / "aniso.npz", **out)
print("wrote", HERE / "aniso.npz")
