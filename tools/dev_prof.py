"""Profiling driver: one rcontrib call on the synthetic office (for ncu)."""
import os, sys, time
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from pyradiance_b200 import _lib, scenegen
TMP = Path(os.environ.get("RB_TMP", "/tmp/rbt")); TMP.mkdir(parents=True, exist_ok=True)
npolys = int(os.environ.get("NPOLY", 100000)); nsens = int(os.environ.get("NSENS", 512))
ab = int(os.environ.get("AB", 3)); ad = int(os.environ.get("AD", 4096)); reps = int(os.environ.get("REPS", 1))
curved = os.environ.get("CURVED", "1") != "0"
rad = TMP / f"off{npolys}{"" if curved else "p"}.rad"; octf = TMP / f"off{npolys}{"" if curved else "p"}.oct"
if not octf.exists():
    scenegen.write_office(rad, npolys=npolys, seed=1234, curved=curved); scenegen.build_octree(rad, octf)
sens = scenegen.office_sensors(nsens)
ctx = _lib.Context(0, _lib.RB_PROGRAM_RCONTRIB); ctx.load_octree(octf)
ctx.set_options(["-ab", str(ab), "-ad", str(ad), "-lw", f"{1.0/ad:.3e}"])
ctx.cal_load("reinhartb.cal"); p = "MF=1,rNx=0,rNy=0,rNz=-1,Ux=0,Uy=1,Uz=0,RHS=+1"; ctx.cal_set(p)
ctx.add_modifier("skyglow", p, "rbin", 145)
for r in range(reps):
    ctx.reset_stats(); t = time.time()
    m = ctx.rcontrib(sens, flags=_lib.RB_IRRAD_RCONTRIB); dt = time.time() - t
    st = ctx.stats()
    print(f"rep {r}: {nsens} sensors {dt:.3f}s rays {st['nrays']} {st['nrays']/dt/1e6:.1f} Mrays/s wall; k_trace {st["wave_ms"]:.1f} ms shade {st["shade_ms"]:.1f} ms -> {st['nrays']/st['wave_ms']/1e3:.1f} Mrays/s; nodes/ray {st['nodes']/st['nrays']:.1f} leafents/ray {st['leafents']/st['nrays']:.1f} prims/ray {st['prims']/st['nrays']:.1f} sum {m.sum():.3f}")
