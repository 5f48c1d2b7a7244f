"""Where the wall time of a small rtrace call goes (BASELINE configs[0]): context, load, trace, close."""
import sys, time
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from pyradiance_b200 import _lib
import pyradiance_b200 as pr
octf = ROOT / "tests" / "golden" / "trace.oct"
gx, gy = np.meshgrid(np.linspace(1, 39, 100), np.linspace(2, 45, 100))
grid = np.stack([gx.ravel(), gy.ravel(), np.full(10000, 2.5), np.zeros(10000), np.zeros(10000), np.ones(10000)], 1)
for rep in range(4):
    t0 = time.perf_counter(); ctx = _lib.Context(0)
    t1 = time.perf_counter(); ctx.load_octree(octf)
    t2 = time.perf_counter(); ctx.set_options(["-ab", "0", "-dt", "0", "-dj", "0", "-dc", "1"])
    v, _ = ctx.rtrace(grid, flags=_lib.RB_IRRAD_RTRACE, want_results=False)
    t3 = time.perf_counter(); v2, _ = ctx.rtrace(grid, flags=_lib.RB_IRRAD_RTRACE, want_results=False)
    t4 = time.perf_counter(); st = ctx.stats(); ctx.close()
    t5 = time.perf_counter()
    print(f"rep {rep}: create {1e3*(t1-t0):.2f} load {1e3*(t2-t1):.2f} first trace {1e3*(t3-t2):.2f} second trace {1e3*(t4-t3):.2f} "
          f"close {1e3*(t5-t4):.2f} ms; kernel_ms {st['kernel_ms']:.3f} launches {st['launches']}")
raw = grid.tobytes()
for rep in range(4):
    t = time.perf_counter()
    out = pr.rtrace(raw, str(octf), header=False, inform="d", outform="d", params=["-I", "-ab", "0", "-dt", "0", "-dj", "0", "-dc", "1"])
    print(f"pr.rtrace call {1e3*(time.perf_counter()-t):.2f} ms")
