"""Developer helper: where the smooth-mesh golden values differ."""
import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import pyradiance_b200 as pr
golden = ROOT / "tests" / "golden"
g = np.load(golden / "smooth.npz")
rays = g["rays"]
out = pr.rtrace(rays.tobytes(), str(golden / "smooth" / "smoothroom.oct"), header=False, inform="d", outform="a",
                outspec="vNnLsm", params=[str(a) for a in g["args"]]).decode()
rows = [ln.split("\t") for ln in out.splitlines()]
val = np.array([[float(x) for x in r[0:3]] for r in rows])
bad = ~np.isclose(val, g["value"], rtol=2e-5, atol=1e-7).all(1)
for i in np.flatnonzero(bad):
    print(i, g["surf"][i], g["mod"][i], "ours", val[i], "ref", g["value"][i], "ratio", val[i, 1] / max(g["value"][i, 1], 1e-30),
          "rod", -(g["fnorm"][i] * rays[i, 3:]).sum(), "pn", g["pnorm"][i], "fn", g["fnorm"][i])
