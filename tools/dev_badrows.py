"""Developer helper: which rays of the smooth-mesh / dielectric / curved-text goldens differ from the reference's
value by more than 1e-5, and by how much (the tests tolerate a handful; this lists them)."""
import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import pyradiance_b200 as pr
G = ROOT / "tests" / "golden"

def run(tag, npz, octf, spec, vcol, extra=()):
    g = np.load(G / npz)
    rays = g["rays"] if "rays" in g else g["curved_rays"]
    args = [str(a) for a in g["args"]]
    out = pr.rtrace(rays.tobytes(), str(octf), header=False, inform="d", outform="a", outspec=spec, params=args).decode()
    rows = [ln.split("\t") for ln in out.splitlines()]
    val = np.array([[float(x) for x in r[vcol:vcol + 3]] for r in rows])
    want = g["value"] if "value" in g else g["curved_value"]
    rel = np.abs(val - want) / (np.abs(want) + 1e-9)
    bad = np.flatnonzero(~np.isclose(val, want, rtol=1e-5, atol=1e-9).all(1))
    print(f"== {tag}: {len(bad)} of {len(rays)} rays beyond 1e-5")
    for i in bad:
        print(f"   ray {i}: org {np.round(rays[i,:3],4)} dir {np.round(rays[i,3:],4)} got {val[i]} want {want[i]} rel {rel[i].max():.2e} cols {rows[i][vcol+3:]}")

run("dielectric", "dielectric.npz", G / "dielectric" / "diel.oct", "vLsm", 0)
try:
    run("smooth", "smooth.npz", G / "smooth" / "smoothroom.oct", "vNnLsm", 0)
except Exception as e:
    print("smooth:", e)
