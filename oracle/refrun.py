"""TEST INFRASTRUCTURE -- runs the compiled reference programs in oracle/_ref/.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module.  The binaries are the UNMODIFIED
reference rtrace / rcontrib / oconv built by oracle/Makefile from
/root/reference; they travel to the GPU box inside oracle/_ref/ (git-ignored).
"""
from __future__ import annotations

import os
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
REF = HERE / "_ref"
BIN = REF / "bin"
LIB = REF / "lib"


def available() -> bool:
    return (BIN / "rtrace").exists() and (BIN / "rcontrib").exists() and (LIB / "rayinit.cal").exists()


def _env():
    env = dict(os.environ)
    extra = os.environ.get("RB_RAYPATH_EXTRA", "")       # e.g. the directory of nested octrees / meshes
    env["RAYPATH"] = f".:{LIB}" + (f":{extra}" if extra else "")
    return env


def run(prog, args, stdin_bytes=b"", check=True, timeout=None):
    """Run oracle/_ref/bin/<prog>; returns stdout bytes (stderr raised on failure)."""
    cmd = [str(BIN / prog)] + [str(a) for a in args]
    r = subprocess.run(cmd, input=stdin_bytes, capture_output=True, env=_env(), timeout=timeout)
    if check and r.returncode != 0:
        raise RuntimeError(f"{prog} failed ({r.returncode}): {r.stderr.decode(errors='replace')}")
    return r.stdout


def rtrace(octree, rays, args, outform="a"):
    """rays: float64 [n,6].  Returns raw stdout (ascii) or an ndarray ('d'/'f')."""
    rays = np.ascontiguousarray(rays, dtype=np.float64)
    out = run("rtrace", ["-h", f"-fd{outform}"] + list(args) + [octree], rays.tobytes())
    if outform == "a":
        return out.decode()
    return np.frombuffer(out, dtype=np.float64 if outform == "d" else np.float32)


def rcontrib(octree, rays, args, nproc=1, outform="d", timeout=None):
    """Returns the coefficient matrix as float64/float32 [nrecords, ncols*3]."""
    rays = np.ascontiguousarray(rays, dtype=np.float64)
    a = ["-n", str(nproc), "-h", f"-fd{outform}"] + list(args) + [octree]
    out = run("rcontrib", a, rays.tobytes(), timeout=timeout)
    return np.frombuffer(out, dtype=np.float64 if outform == "d" else np.float32)


def oconv(rad_files, oct_path, frozen=True):
    args = (["-f"] if frozen else []) + [str(f) for f in rad_files]
    data = run("oconv", args)
    Path(oct_path).write_bytes(data)
