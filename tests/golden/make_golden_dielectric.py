"""Golden fixture for the refracting materials dielectric / interface (rt/dielectric.c) and the
path extinction they switch on (rayparticipate(), rt/raytrace.c:259-295), SURVEY 8f row f4.

TEST INFRASTRUCTURE.  Run in the build container (needs oracle/_ref):

    python tests/golden/make_golden_dielectric.py

Writes tests/golden/dielectric/{diel.rad,diel.oct} and tests/golden/dielectric.npz.  Scene: a
tinted dielectric slab (a closed box: rays enter, are attenuated per unit length inside, leave
or are totally reflected), a clear dielectric ball (a lens), a box of `interface` material
(water in glass: different index and tint on either side) standing on a checker of three
coloured plastic tiles, a distant sun, a local lamp and a glow sky.  Deterministic settings
(-ab 0 -dt 0 -dj 0 -dc 1 -st 1 -lr 8 -lw 1e-3 -av .05 .05 .05; -lr > 0 so no Russian roulette):
rtrace values, distances and names of 3000 view rays aimed at the three bodies, plus -I
sensors under them (shadow rays refract and are attenuated, never reflected), and rcontrib
coefficients of the three emitters for 600 of the rays (they carry the path extinction too).
"""
import os
import subprocess
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent.parent))
from oracle import refrun  # noqa: E402

S = HERE / "dielectric"
S.mkdir(exist_ok=True)
env = dict(os.environ, RAYPATH=f".:{refrun.LIB}")


def box(mod, name, lo, hi):
    x0, y0, z0 = lo; x1, y1, z1 = hi
    f = [((x0, y0, z0), (x0, y1, z0), (x1, y1, z0), (x1, y0, z0)), ((x0, y0, z1), (x1, y0, z1), (x1, y1, z1), (x0, y1, z1)),
         ((x0, y0, z0), (x1, y0, z0), (x1, y0, z1), (x0, y0, z1)), ((x0, y1, z0), (x0, y1, z1), (x1, y1, z1), (x1, y1, z0)),
         ((x0, y0, z0), (x0, y0, z1), (x0, y1, z1), (x0, y1, z0)), ((x1, y0, z0), (x1, y1, z0), (x1, y1, z1), (x1, y0, z1))]
    out = ""
    for i, q in enumerate(f):
        out += f"{mod} polygon {name}.{i}\n0\n0\n12 " + "  ".join(" ".join(f"{c:g}" for c in v) for v in q) + "\n\n"
    return out


rad = """void plastic tile_r
0
0
5 .6 .2 .2 0 0

void plastic tile_g
0
0
5 .2 .6 .2 0 0

void plastic tile_b
0
0
5 .2 .2 .6 .02 .05

tile_r polygon floor_r
0
0
12 -2 -2 0  3 -2 0  3 8 0  -2 8 0

tile_g polygon floor_g
0
0
12 3 -2 0  8 -2 0  8 8 0  3 8 0

tile_b polygon floor_b
0
0
12 8 -2 0  14 -2 0  14 8 0  8 8 0

void dielectric tinted
0
0
5 .6 .8 .7 1.52 0

void dielectric clear
0
0
5 .98 .98 .98 1.33 0

void interface water_in_glass
0
0
8 .8 .9 .95 1.33  .9 .7 .9 1.52

void light sunl
0
0
3 6000 6000 5500

sunl source sun
0
0
4 .25 -.35 .9 1.2

void light lampl
0
0
3 50 45 35

lampl polygon lamp
0
0
12 4 2 5  7 2 5  7 4 5  4 4 5

void glow skyg
0
0
4 .7 .8 1.1 0

skyg source sky
0
0
4 0 0 1 180

clear sphere lens
0
0
4 6 3 1.6 .9

"""
rad += box("tinted", "slab", (0, 1, .5), (2.5, 5, 1.1))
rad += box("water_in_glass", "tank", (9, 1, .3), (11.5, 4, 1.8))
(S / "diel.rad").write_text(rad)
r = subprocess.run([str(refrun.BIN / "oconv"), "-f", "diel.rad"], cwd=S, env=env, capture_output=True)
assert r.returncode == 0, r.stderr.decode()
(S / "diel.oct").write_bytes(r.stdout)

rng = np.random.default_rng(17)
n = 3000
which = rng.integers(0, 3, n)
ctr = np.array([[1.25, 3, .8], [6, 3, 1.6], [10.25, 2.5, 1.05]])[which]
tgt = ctr + rng.normal(size=(n, 3)) * np.array([[.9, 1.4, .25], [.5, .5, .5], [.9, 1.0, .6]])[which]
org = tgt + np.stack([rng.uniform(-4, 4, n), rng.uniform(-4, 4, n), rng.uniform(.5, 4, n)], 1)
d = tgt - org
d /= np.linalg.norm(d, axis=1, keepdims=True)
rays = np.concatenate([org, d], 1)
det = ["-ab", "0", "-dt", "0", "-dj", "0", "-dc", "1", "-st", "1", "-lr", "8", "-lw", "1e-3", "-av", ".05", ".05", ".05"]
txt = subprocess.run([str(refrun.BIN / "rtrace"), "-h", "-fda", "-ovLsm"] + det + ["diel.oct"], cwd=S, env=env,
                     capture_output=True, input=rays.tobytes())
assert txt.returncode == 0, txt.stderr.decode()
rows = [ln.split("\t") for ln in txt.stdout.decode().splitlines()]
out = {"rays": rays, "args": np.array(det),
       "value": np.array([[float(x) for x in q[0:3]] for q in rows]), "dist": np.array([float(q[3]) for q in rows]),
       "surf": np.array([q[4] for q in rows]), "mod": np.array([q[5] for q in rows])}
print({k: int((out["mod"] == k).sum()) for k in ("tinted", "clear", "water_in_glass", "tile_r", "tile_g", "tile_b")})
sens = np.array([[x, y, .01, 0, 0, 1] for x in np.linspace(-.5, 12.5, 27) for y in (1.5, 2.5, 3, 3.5, 4.5)], dtype=float)
out["sensors"] = sens
out["irrad"] = refrun.rtrace(S / "diel.oct", sens, ["-I"] + det, outform="d").reshape(-1, 3)
# Russian roulette on (-lr -10 default of rcontrib; here rtrace with -lr 0): means over repetitions
st = ["-ab", "0", "-dt", "0", "-dj", "0", "-dc", "1", "-st", "1", "-lr", "-10", "-lw", "2e-2", "-av", ".05", ".05", ".05"]
pick = np.flatnonzero(np.isin(out["mod"], ["tinted", "clear", "water_in_glass"]))[:150]
reps = 1200
v = refrun.rtrace(S / "diel.oct", np.tile(rays[pick], (reps, 1)), st, outform="d").reshape(reps, len(pick), 3)
out["rr_pick"], out["rr_args"] = pick, np.array(st)
out["rr_mean"], out["rr_sem"] = v.mean(0), v.std(0, ddof=1) / np.sqrt(reps)
# rcontrib coefficients per emitter: raycontrib() carries the extinction summed over the chain (raytrace.c:407-442)
rc = ["-ab", "0", "-dt", "0", "-dj", "0", "-dc", "1", "-st", "1", "-lr", "8", "-lw", "1e-3"]
out["rc_args"] = np.array(rc)
out["rc"] = refrun.rcontrib(S / "diel.oct", rays[:600], rc + ["-m", "skyg", "-m", "lampl", "-m", "sunl"]).reshape(600, 3, 3)
np.savez_compressed(HERE / "dielectric.npz", **out)
print("irradiance range", out["irrad"].min(), out["irrad"].max(), "wrote", HERE / "dielectric.npz")
