import os, sys, time
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from pyradiance_b200 import _lib, scenegen
TMP = Path(os.environ.get("RB_TMP", "/tmp/rbt"))
octf = TMP / "off100000.oct"
sens = scenegen.office_sensors(int(os.environ.get("NSENS", 20000)))
ctx = _lib.Context(0, _lib.RB_PROGRAM_RCONTRIB); ctx.load_octree(octf)
ctx.set_options(["-ab", "3", "-ad", "4096", "-lw", f"{1.0/4096:.3e}"])
ctx.cal_load("reinhartb.cal"); p = "MF=1,rNx=0,rNy=0,rNz=-1,Ux=0,Uy=1,Uz=0,RHS=+1"; ctx.cal_set(p)
ctx.add_modifier("skyglow", p, "rbin", 145)
for r in range(3):
    ctx.reset_stats(); t = time.time()
    m = ctx.rcontrib(sens, flags=_lib.RB_IRRAD_RCONTRIB); dt = time.time() - t
    st = ctx.stats()
    print(f"wall {dt*1e3:.1f} ms", {k: (round(v, 1) if isinstance(v, float) else v) for k, v in st.items() if k in ("kernel_ms", "wave_ms", "shade_ms", "launches", "waves", "batches", "retries")})
