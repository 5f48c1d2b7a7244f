"""Golden fixture for the rmtxop mirror (pyradiance_b200/mtx.py), SURVEY 8f row f2.

TEST INFRASTRUCTURE.  Run in the build container (needs oracle/_ref/bin/rmtxop, the unmodified
reference built by oracle/Makefile):

    python tests/golden/make_golden_rmtxop.py

Writes tests/golden/rmtxop/{A,B,C,D,V,S}.mtx (seeded small matrices in ascii / double / float form,
1- and 3-component) and tests/golden/rmtxop.npz: for every command line in CASES the bytes the
reference printed.  Cases without a matrix product are reproduced byte for byte by the mirror
(float storage and the reference's operation order), except element-wise division, which the
reference's -ffast-math build vectorises with a reciprocal approximation (last-bit differences);
products run on the GPU in fp32 and are compared within 1e-5.
"""
import subprocess
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent.parent))
from oracle import refrun  # noqa: E402

S = HERE / "rmtxop"
S.mkdir(exist_ok=True)
rng = np.random.default_rng(3)


def write(name, m, fmt):
    nr, nc, k = m.shape
    h = f"#?RADIANCE\nNROWS={nr}\nNCOLS={nc}\nNCOMP={k}\n"
    if fmt == "a":
        body = "\n".join(" ".join("%.9e" % v for v in r.ravel()) for r in m) + "\n"
        (S / name).write_bytes((h + "FORMAT=ascii\n\n" + body).encode())
    else:
        h += "BigEndian=0\nFORMAT=" + ("float" if fmt == "f" else "double") + "\n\n"
        (S / name).write_bytes(h.encode() + m.astype("<f4" if fmt == "f" else "<f8").tobytes())


write("A.mtx", rng.random((7, 5, 3)), "a")
write("B.mtx", rng.random((5, 4, 3)), "d")
write("C.mtx", rng.random((7, 5, 3)), "f")
write("D.mtx", rng.random((7, 5, 1)) + .1, "a")
write("V.mtx", rng.random((40, 145, 3)) * 1e-3, "f")          # a view / daylight-coefficient block
write("S.mtx", rng.random((145, 24, 3)) * 1e4, "a")           # a sky matrix: 24 time steps

CASES = [
    ["-fa", "A.mtx"], ["-t", "A.mtx"], ["-s", "2.5", "A.mtx"], ["-s", "1", "2", "3", "A.mtx"],
    ["-c", "0.2", "0.7", "0.1", "A.mtx"], ["-c", "XYZ", "A.mtx"], ["-c", "RGB", "-t", "A.mtx"], ["-c", "Y", "A.mtx"],
    ["-c", "y", "A.mtx"], ["-c", "A", "A.mtx"], ["-fd", "A.mtx", "+", "C.mtx"], ["A.mtx", "+", "-s", ".5", "C.mtx"],
    ["A.mtx", "*", "C.mtx"], ["A.mtx", "/", "D.mtx"], ["A.mtx", "*", "D.mtx"], ["-ff", "C.mtx"],
    ["A.mtx", "+", "C.mtx", "-c", "0.3", "0.6", "0.1"], ["-s", "2", "-c", "1", "0", "0", "0", "1", "0", "A.mtx"],
    ["-C", "Y", "A.mtx", "+", "C.mtx"], ["-fa", "A.mtx", "/", "C.mtx"], ["-fa", "B.mtx"],
    # products (GPU)
    ["A.mtx", ".", "B.mtx"], ["A.mtx", "B.mtx"], ["-t", "B.mtx", ".", "-t", "A.mtx"], ["-fd", "V.mtx", "S.mtx"],
    ["V.mtx", "S.mtx", "-c", "47.4", "119.9", "11.6"], ["-s", "0.5", "C.mtx", ".", "B.mtx", ".", "-t", "B.mtx", "+", "A.mtx"],
    ["-c", "Y", "V.mtx", ".", "-c", "Y", "S.mtx"],
]
out = {"cases": np.array(["\x1f".join(c) for c in CASES])}
for i, c in enumerate(CASES):
    r = subprocess.run([str(refrun.BIN / "rmtxop")] + c, cwd=S, capture_output=True)
    assert r.returncode == 0, (c, r.stderr.decode())
    out[f"out{i}"] = np.frombuffer(r.stdout, dtype=np.uint8)
np.savez_compressed(HERE / "rmtxop.npz", **out)
print("wrote", HERE / "rmtxop.npz", len(CASES), "cases")
