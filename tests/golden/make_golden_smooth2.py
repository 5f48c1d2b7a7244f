"""Second golden fixture for smooth (vertex-normal) mesh triangles: the materials added after the first
one -- plastic2, metal2, trans2 (rt/aniso.c) and dielectric (rt/dielectric.c) -- all of which look at
RAY.pert (perturbed normal, bent transmission with its penetration guards, the "Phong" exemption).

TEST INFRASTRUCTURE.  Run in the build container after make_golden_smooth.py (reuses its smooth.obj,
smoothroom.rad and rays; needs oracle/_ref):

    python tests/golden/make_golden_smooth2.py

Writes tests/golden/smooth/{smoothmats2.rad,smooth2.rtm,smoothroom2.rad,smoothroom2.oct} and
tests/golden/smooth2.npz (the reference rtrace's value, distance, names, -oN and -on for the rays of
smooth.npz).  Deterministic settings: -ab 0 -dt 0 -dj 0 -dc 1 -lr 6 -lw 1e-4 -st 1.
"""
import os
import subprocess
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent.parent))
from oracle import refrun  # noqa: E402

S = HERE / "smooth"
env = dict(os.environ, RAYPATH=f".:{refrun.LIB}")


def sh(cmd, out=None, stdin=None):
    r = subprocess.run(cmd, cwd=S, env=env, capture_output=True, input=stdin)
    assert r.returncode == 0, r.stderr.decode()
    if out:
        (S / out).write_bytes(r.stdout)
    return r.stdout


(S / "smoothmats2.rad").write_text("""void plastic2 sm_plastic
4 1 0 0 .
0
6 .6 .5 .4 .06 .1 .3

void metal2 sm_metal
4 0 1 .3 .
0
6 .7 .6 .3 .8 .25 .07

void dielectric sm_glass
0
0
5 .8 .9 .85 1.5 0

void trans2 sm_trans
4 1 1 0 .
0
8 .7 .7 .7 .05 .15 .1 .6 .7

void trans2 Phong
4 1 1 0 .
0
8 .7 .7 .7 .05 .15 .1 .6 .7
""")
room = (S / "smoothroom.rad").read_text().replace("smooth.rtm", "smooth2.rtm")
(S / "smoothroom2.rad").write_text(room)
sh([str(refrun.BIN / "obj2mesh"), "-a", "smoothmats2.rad", "smooth.obj", "smooth2.rtm"])
sh([str(refrun.BIN / "oconv"), "-f", "smoothroom2.rad"], "smoothroom2.oct")
g = np.load(HERE / "smooth.npz")
rays = g["rays"]
args = [str(a) for a in g["args"]] + ["-st", "1"]
out = sh([str(refrun.BIN / "rtrace"), "-h", "-fda"] + args + ["-ovNnLsm", "smoothroom2.oct"], stdin=rays.tobytes()).decode()
rows = [ln.split("\t") for ln in out.splitlines()]
assert len(rows) == len(rays)
val = np.array([[float(x) for x in r[0:3]] for r in rows])
fn = np.array([[float(x) for x in r[3:6]] for r in rows])
pn = np.array([[float(x) for x in r[6:9]] for r in rows])
dist = np.array([float(r[9]) for r in rows])
surf = np.array([r[10] for r in rows])
mod = np.array([r[11] for r in rows])
smooth = (np.abs(np.abs(pn) - np.abs(fn)).max(1) > 1e-6)
print({m: int(((mod == m) & smooth).sum()) for m in sorted(set(mod))})
np.savez_compressed(HERE / "smooth2.npz", args=np.array(args), value=val, pnorm=pn, fnorm=fn, dist=dist, surf=surf, mod=mod)
