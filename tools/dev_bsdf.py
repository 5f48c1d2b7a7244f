"""Developer helper (GPU box): BSDF / aBSDF golden, per-modifier deviation from the reference."""
import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import pyradiance_b200 as pr
from pyradiance_b200 import _lib
G = ROOT / "tests" / "golden"
g = np.load(G / "bsdfmat.npz")
rays = g["rays"]
for tag, octf in (("", "bsdfmat.oct"), ("lamp_", "bsdflamp.oct")):
    out = pr.rtrace(rays.tobytes(), str(G / "bsdfmat" / octf), header=False, inform="d", outform="a", outspec="vLsm",
                    params=[str(a) for a in g["args"]]).decode()
    rows = [ln.split("\t") for ln in out.splitlines()]
    surf = np.array([r[4] for r in rows]); mod = np.array([r[5] for r in rows])
    print(tag, "names equal:", (surf == g[tag + "surf"]).all(), (mod == g[tag + "mod"]).all())
    val = np.array([[float(x) for x in r[0:3]] for r in rows])
    ref = g[tag + "value"]
    rel = np.abs(val - ref).max(1) / np.maximum(np.abs(ref).max(1), 1e-9)
    for m in np.unique(mod):
        k = mod == m
        below = rays[k, 5] > 0
        print(f"  {m:8s} n={k.sum():4d} max rel {rel[k].max():.3e}  bad(>1e-5) {(rel[k] > 1e-5).sum():4d}  (from below: {(rel[k][below] > 1e-5).sum()} of {below.sum()})")
    bad = np.flatnonzero(rel > 1e-5)[:6]
    for i in bad: print("   ray", i, mod[i], rays[i].round(3), "got", val[i], "want", ref[i])
    ctx = _lib.Context(0)
    ctx.load_octree(G / "bsdfmat" / octf)
    ctx.set_options([str(a) for a in g["args"]])
    v, _ = ctx.rtrace(g["sensors"], flags=_lib.RB_IRRAD_RTRACE)
    ri = np.abs(v - g[tag + "irrad"]).max(1) / np.maximum(np.abs(g[tag + "irrad"]).max(1), 1e-9)
    print("  irrad max rel", ri.max(), "bad", np.flatnonzero(ri > 1e-5))
    for i in np.flatnonzero(ri > 1e-5)[:8]: print("    sensor", g["sensors"][i], "got", v[i], "want", g[tag + "irrad"][i])
ctx = _lib.Context(0)
ctx.load_octree(G / "bsdfmat" / "bsdfmat.oct")
pick, reps = g["st_pick"], 1200
ctx.set_options([str(a) for a in g["st_args"]])
v, _ = ctx.rtrace(np.tile(rays[pick], (reps, 1)))
v = v.reshape(reps, len(pick), 3)
sem = np.sqrt(v.var(0, ddof=1) / reps + g["st_sem"] ** 2)
z = np.abs(v.mean(0) - g["st_mean"]) / (sem + 1e-6 * g["st_mean"] + 1e-12)
print("stochastic view rays: max z", z.max(), "n(z>5)", (z.max(1) > 5).sum(), "of", len(pick))
mod = g["mod"][pick]
for i in np.flatnonzero(z.max(1) > 5)[:10]: print("   ", mod[i], rays[pick[i]].round(3), "got", v.mean(0)[i], "want", g["st_mean"][i], "sem", sem[i])
ctx = _lib.Context(0)
ctx.load_octree(G / "bsdfmat" / "bsdfmat.oct")
ctx.set_options([str(a) for a in g["ab1_args"]])
s2 = g["ab1_sensors"]; reps2 = 150
v, _ = ctx.rtrace(np.tile(s2, (reps2, 1)), flags=_lib.RB_IRRAD_RTRACE)
v = v.reshape(reps2, len(s2), 3)
sem = np.sqrt(v.var(0, ddof=1) / reps2 + g["ab1_sem"] ** 2)
z = np.abs(v.mean(0) - g["ab1_mean"]) / (sem + 1e-12)
print("ab1 sensors: max z", z.max(), "rel dev max", (np.abs(v.mean(0) - g["ab1_mean"]) / g["ab1_mean"]).max())
for i in np.flatnonzero(z.max(1) > 5)[:10]: print("   ", s2[i], "got", v.mean(0)[i], "want", g["ab1_mean"][i], "sem", sem[i])
rc = _lib.Context(0, _lib.RB_PROGRAM_RCONTRIB)
rc.load_octree(G / "bsdfmat" / "bsdfmat.oct")
rc.set_options([str(a) for a in g["rc_args"]])
for m in ("skyg", "sunl", "gndg"): rc.add_modifier(m, "", "0", 1)
m = rc.rcontrib(np.tile(s2, (reps2, 1)), flags=_lib.RB_IRRAD_RCONTRIB, dtype=np.float64).reshape(reps2, len(s2), 3, 3)
sem = np.sqrt(m.var(0, ddof=1) / reps2 + g["rc_sem"] ** 2)
z = np.abs(m.mean(0) - g["rc_mean"]) / (sem + 0.004 * g["rc_mean"] + 1e-12)
print("rcontrib: max z", z.max())
for i in np.flatnonzero(z.reshape(len(s2), -1).max(1) > 5)[:10]: print("   ", s2[i], "got", m.mean(0)[i].ravel(), "want", g["rc_mean"][i].ravel())
