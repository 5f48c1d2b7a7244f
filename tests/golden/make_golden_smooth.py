"""Golden fixture for smooth (vertex-normal) mesh triangles, SURVEY 8a row a10.

TEST INFRASTRUCTURE.  Run in the build container (needs oracle/_ref, i.e. the
unmodified reference obj2mesh / oconv / rtrace built by oracle/Makefile):

    python tests/golden/make_golden_smooth.py

Writes tests/golden/smooth/{smooth.obj,smoothmats.rad,smooth.rtm,smoothroom.rad,
smoothroom.oct} and tests/golden/smooth.npz (rays + the reference rtrace's
answers: value, unperturbed normal -oN, perturbed normal -on, distance, surface and
modifier names).  The mesh is a height field with analytic vertex normals, a
strip of faces WITHOUT normals (those stay flat), five materials (plastic with
a highlight, pure-specular metal, glass, pure-specular trans, and a trans
named "Phong", which the reference exempts from perturbed transmission,
rt/rtotypes.h:13), placed three times: as is, rotated + scaled, mirrored with
a modifier override.  Deterministic settings: -ab 0 -dt 0 -dj 0 -dc 1.
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
S.mkdir(exist_ok=True)
env = dict(os.environ, RAYPATH=f".:{refrun.LIB}")


def sh(cmd, out=None, stdin=None):
    r = subprocess.run(cmd, cwd=S, env=env, capture_output=True, input=stdin)
    assert r.returncode == 0, r.stderr.decode()
    if out:
        (S / out).write_bytes(r.stdout)
    return r.stdout


def height(x, y):
    return 0.25 * np.sin(1.7 * x) * np.cos(1.3 * y) + 0.1 * np.sin(3.1 * x + 0.5 * y)


def normal(x, y):
    dzdx = 0.25 * 1.7 * np.cos(1.7 * x) * np.cos(1.3 * y) + 0.1 * 3.1 * np.cos(3.1 * x + 0.5 * y)
    dzdy = -0.25 * 1.3 * np.sin(1.7 * x) * np.sin(1.3 * y) + 0.1 * 0.5 * np.cos(3.1 * x + 0.5 * y)
    n = np.array([-dzdx, -dzdy, 1.0])
    return n / np.linalg.norm(n)


N = 13
xs = np.linspace(0, 3, N)
ys = np.linspace(0, 3, N)
mats = ["sm_plastic", "sm_metal", "sm_glass", "sm_trans", "Phong"]
with open(S / "smooth.obj", "w") as f:
    for i in range(N):
        for j in range(N):
            f.write(f"v {xs[i]:.6f} {ys[j]:.6f} {height(xs[i], ys[j]):.6f}\n")
    for i in range(N):
        for j in range(N):
            n = normal(xs[i], ys[j])
            f.write(f"vn {n[0]:.6f} {n[1]:.6f} {n[2]:.6f}\n")
    vid = lambda i, j: i * N + j + 1
    for i in range(N - 1):
        f.write(f"usemtl {mats[(i * 5) // (N - 1)]}\n")
        for j in range(N - 1):
            a, b, c, d = vid(i, j), vid(i + 1, j), vid(i + 1, j + 1), vid(i, j + 1)
            if j == 5:                     # one strip without normals: flat triangles among smooth ones
                f.write(f"f {a} {b} {c}\nf {a} {c} {d}\n")
            else:
                f.write(f"f {a}//{a} {b}//{b} {c}//{c}\nf {a}//{a} {c}//{c} {d}//{d}\n")

(S / "smoothmats.rad").write_text("""void plastic sm_plastic
0
0
5 .6 .6 .6 .05 .08

void metal sm_metal
0
0
5 .7 .6 .3 .9 0

void glass sm_glass
0
0
3 .8 .8 .8

void trans sm_trans
0
0
7 .7 .7 .7 .1 0 .6 .9

void trans Phong
0
0
7 .7 .7 .7 .1 0 .6 .9
""")

(S / "smoothroom.rad").write_text("""void plastic red
0
0
5 .6 .2 .2 0 0

void plastic blue
0
0
5 .2 .2 .7 0 0

void plastic green
0
0
5 .2 .6 .2 .03 .1

red polygon floor
0
0
12 -6 -6 -1  6 -6 -1  6 6 -1  -6 6 -1

blue polygon wall
0
0
12 -6 6 -1  6 6 -1  6 6 5  -6 6 5

void light sunl
0
0
3 8000 8000 7000

sunl source sun
0
0
4 .35 -.45 .82 2

void glow skyg
0
0
4 .9 .9 1.3 0

skyg source sky
0
0
4 0 0 1 180

void mesh sm_a
1 smooth.rtm
0
0

void mesh sm_b
9 smooth.rtm -rx 35 -s 1.4 -t -5 -4 0.8
0
0

green mesh sm_c
6 smooth.rtm -mx -t -1 -4.5 1.5
0
0
""")

sh([str(refrun.BIN / "obj2mesh"), "-a", "smoothmats.rad", "smooth.obj", "smooth.rtm"])
sh([str(refrun.BIN / "oconv"), "-f", "smoothroom.rad"], "smoothroom.oct")

rng = np.random.default_rng(11)
n = 3000
# targets on / around the three placements, origins above and below them
tg = np.concatenate([
    rng.uniform((0, 0, -0.2), (3, 3, 0.4), size=(n // 3, 3)),
    rng.uniform((-5, -4, 0.2), (-0.8, 0.5, 3.2), size=(n // 3, 3)),
    rng.uniform((-4, -4.5, 1.2), (-1, -1.5, 1.9), size=(n - 2 * (n // 3), 3)),
])
org = tg + rng.normal(size=(n, 3)) * (1.5, 1.5, 0.4) + np.where(rng.random(n) < 0.7, 1, -1)[:, None] * (0, 0, 2.5)
d = tg - org
d /= np.linalg.norm(d, axis=1, keepdims=True)
rays = np.concatenate([org, d], axis=1)
args = ["-ab", "0", "-dt", "0", "-dj", "0", "-dc", "1", "-lr", "6", "-lw", "1e-4"]
out = sh([str(refrun.BIN / "rtrace"), "-h", "-fda"] + args + ["-ovNnLsm", "smoothroom.oct"], stdin=rays.tobytes()).decode()
rows = [ln.split("\t") for ln in out.splitlines()]
assert len(rows) == n
val = np.array([[float(x) for x in r[0:3]] for r in rows])
fn = np.array([[float(x) for x in r[3:6]] for r in rows])      # -oN: unperturbed, flips undone
pn = np.array([[float(x) for x in r[6:9]] for r in rows])      # -on: perturbed, as shading left it
dist = np.array([float(r[9]) for r in rows])
surf = np.array([r[10] for r in rows])
mod = np.array([r[11] for r in rows])
smooth = (np.abs(np.abs(pn) - np.abs(fn)).max(1) > 1e-6)
print(f"{n} rays: {np.sum(surf == 'M-Tri')} on mesh triangles, {np.sum(surf == 'sm_c')} on the overridden mesh, "
      f"{smooth.sum()} with a perturbed normal; modifiers hit: {sorted(set(mod))}")
np.savez_compressed(HERE / "smooth.npz", rays=rays, args=np.array(args), value=val, pnorm=pn, fnorm=fn, dist=dist,
                    surf=surf, mod=mod)
