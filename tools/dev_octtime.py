import sys, time, os
sys.path.insert(0, '.')
from pathlib import Path
from pyradiance_b200 import _lib, scenegen
tmp = Path('/tmp/rbt'); tmp.mkdir(exist_ok=True)
n = int(os.environ.get("NPOLY", 1000000))
rad = tmp / f"t{n}.rad"
if not rad.exists(): scenegen.write_office(rad, npolys=n, floors=10 if n >= 500000 else 1, seed=77)
_lib.Context(0)   # CUDA context up
for mode in ("dev", "host", "dev"):
    if mode == "host": os.environ["RB_OCTBUILD_HOST"] = "1"
    else: os.environ.pop("RB_OCTBUILD_HOST", None)
    t = time.time(); _lib.oconv_file(rad, tmp / f"t_{mode}.oct"); print(mode, f"{time.time()-t:.2f} s", flush=True)
