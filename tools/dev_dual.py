"""Developer experiment (GPU box): one context tracing N sensors vs two contexts on the same GPU tracing N/2 each at once."""
import os, sys, time, threading
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from pyradiance_b200 import _lib, scenegen
TMP = Path(os.environ.get("RB_TMP", "/tmp/rbt")); TMP.mkdir(parents=True, exist_ok=True)
npolys = 100000; nsens = int(os.environ.get("NSENS", 20000))
rad = TMP / f"off{npolys}.rad"; octf = TMP / f"off{npolys}.oct"
if not octf.exists():
    scenegen.write_office(rad, npolys=npolys, seed=1234); scenegen.build_octree(rad, octf)
sens = scenegen.office_sensors(nsens)
def mk():
    ctx = _lib.Context(0, _lib.RB_PROGRAM_RCONTRIB); ctx.load_octree(octf)
    ctx.set_options(["-ab", "3", "-ad", "4096", "-lw", f"{1.0/4096:.3e}"])
    ctx.cal_load("reinhartb.cal"); p = "MF=1,rNx=0,rNy=0,rNz=-1,Ux=0,Uy=1,Uz=0,RHS=+1"; ctx.cal_set(p)
    ctx.add_modifier("skyglow", p, "rbin", 145)
    return ctx
K = int(os.environ.get("NCTX", 2))
ctxs = [mk() for _ in range(K)]
out = np.empty((nsens, 145, 3), np.float32)
for rep in range(3):
    t = time.time(); ctxs[0].rcontrib(sens, flags=_lib.RB_IRRAD_RCONTRIB, out=out); t1 = time.time() - t
    ref = out.copy()
    bounds = np.linspace(0, nsens, K + 1).astype(int)
    def work(k):
        a, b = bounds[k], bounds[k + 1]
        ctxs[k].rcontrib(sens[a:b], flags=_lib.RB_IRRAD_RCONTRIB, row_base=int(a), out=out[a:b])
    th = [threading.Thread(target=work, args=(k,)) for k in range(K)]
    t = time.time(); [x.start() for x in th]; [x.join() for x in th]; t2 = time.time() - t
    print(f"rep {rep}: single {t1*1e3:.1f} ms, {K} contexts at once {t2*1e3:.1f} ms, speed-up {t1/t2:.3f}, identical {np.array_equal(ref, out)}")
