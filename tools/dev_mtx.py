"""Developer timing of the matrix consumer (rb_mtx_multiply) at BASELINE configs[1] size:
DC [100000 x 145 x 3] x sky [145 x 8760 x 3] with everything resident in HBM, next to the
reference dctimestep on a row sample."""
import ctypes as C, os, sys, time, subprocess, tempfile
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
from pyradiance_b200 import _lib
from oracle import refrun
nr, ni, nc = int(os.environ.get("NR", 100000)), int(os.environ.get("NI", 145)), int(os.environ.get("NC", 8760))
print("shape", nr, ni, nc)
g = torch.Generator(device="cuda").manual_seed(1)
a = torch.rand((nr, ni, 3), device="cuda", generator=g) ** 6 * 0.05
b = torch.rand((ni, nc, 3), device="cuda", generator=g) ** 3 * 2e4
out = torch.empty((nr, nc, 3), device="cuda", dtype=torch.float32)
ctx = _lib.Context(0)
ms = C.c_double(0)
flags = 1 | 2 | 4
for it in range(4):
    torch.cuda.synchronize(); t = time.time()
    rv = ctx.lib.rb_mtx_multiply(ctx.h, a.data_ptr(), nr, ni, b.data_ptr(), nc, out.data_ptr(), flags, C.byref(ms))
    torch.cuda.synchronize(); dt = time.time() - t
    assert rv == 0
    fl = 2.0 * nr * ni * nc * 3
    print(f"run {it}: kernel {ms.value:.2f} ms wall {dt*1e3:.2f} ms  {fl/ms.value/1e9:.1f} TFLOP/s fp32 (SIMT peak ~74.5), "
          f"out {out.numel()*4/1e9:.2f} GB written at {out.numel()*4/ms.value/1e6:.0f} GB/s")
ref = torch.einsum("rik,ick->rck", a[:256].double(), b.double())
err = ((out[:256].double() - ref).abs() / ref.clamp(min=1e-30)).max().item()
print("max rel err vs float64 on 256 rows:", err)
if refrun.available():
    n = 300
    with tempfile.TemporaryDirectory() as td:
        def wr(p, m):
            h = f"#?RADIANCE\nNROWS={m.shape[0]}\nNCOLS={m.shape[1]}\nNCOMP=3\nBigEndian=0\nFORMAT=float\n\n".encode()
            Path(p).write_bytes(h + m.cpu().numpy().astype(np.float32).tobytes())
        wr(f"{td}/dc.mtx", a[:n]); wr(f"{td}/sky.smx", b)
        t = time.time(); r = subprocess.run([str(refrun.BIN / "dctimestep"), "-of", f"{td}/dc.mtx", f"{td}/sky.smx"], capture_output=True); dt = time.time() - t
        assert r.returncode == 0, r.stderr
        print(f"reference dctimestep: {n} rows in {dt:.2f} s -> {nr} rows ~ {dt*nr/n:.0f} s (1 core)")
