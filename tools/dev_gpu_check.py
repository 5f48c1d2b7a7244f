"""Developer smoke script for the GPU box (not a test, not the bench):
runs the CUDA path on a few inputs and prints diffs against the reference
binaries in oracle/_ref.  Usage: python tools/dev_gpu_check.py [quick]"""
import json
import os
import sys
import time
import traceback
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from pyradiance_b200 import _lib, scenegen  # noqa: E402
from oracle import refrun  # noqa: E402

G = ROOT / "tests" / "golden"
TMP = Path(os.environ.get("RB_TMP", "/tmp/rbt"))
TMP.mkdir(parents=True, exist_ok=True)


def names(ctx, res):
    s = [ctx.object_name(i) if i >= 0 else "*" for i in res["robj"]]
    m = [ctx.object_name(i) if i >= 0 else "*" for i in res["omod"]]
    return s, m


def check_known():
    g = json.load(open(G / "golden.json"))
    ctx = _lib.Context(0)
    ctx.load_octree(G / "trace.oct")
    ctx.set_options(["-ab", "0"])
    k = g["trace_ovposmNL"]
    vals, res = ctx.rtrace(np.array(k["rays"]))
    s, m = names(ctx, res)
    print("known-answer rays:")
    for i in range(len(s)):
        print("  mine:", vals[i], res["rop"][i], s[i], m[i], res["ron"][i], res["rot"][i])
    print("  ref :\n" + k["out"])
    k = g["trace_I_ab0"]
    ctx2 = _lib.Context(0)
    ctx2.load_octree(G / "trace.oct")
    ctx2.set_options(["-ab", "0", "-dt", "0", "-dj", "0", "-dc", "1"])
    vals, _ = ctx2.rtrace(np.array(k["rays"]), flags=_lib.RB_IRRAD_RTRACE, want_results=False)
    print("-I -ab 0 values mine:\n", vals, "\nref:\n" + k["out"])


def check_office(npolys=2000, nrays=200000):
    rad = TMP / f"off{npolys}.rad"
    octf = TMP / f"off{npolys}.oct"
    scenegen.write_office(rad, npolys=npolys, seed=1)
    scenegen.build_octree(rad, octf)
    rays = scenegen.random_rays(nrays)
    ctx = _lib.Context(0)
    t = time.time()
    ctx.load_octree(octf)
    print(f"office{npolys}: load {time.time() - t:.3f}s warnings={ctx.warnings()[:200]!r}")
    ctx.set_options(["-ab", "0"])
    t = time.time()
    _, res = ctx.rtrace(rays, want_values=False)
    dt = time.time() - t
    st = ctx.stats()
    print(f"  traced {nrays} rays in {dt:.3f}s; stats {st}")
    s, m = names(ctx, res)
    t = time.time()
    ref = refrun.rtrace(octf, rays, ["-ab", "0", "-osmL"]).splitlines()
    print(f"  reference rtrace: {time.time() - t:.3f}s")
    bad = 0
    maxrel = 0.0
    for i, line in enumerate(ref):
        f = line.split("\t")
        rs, rm, rl = f[0], f[1], float(f[2])
        if rs != s[i] or rm != m[i]:
            bad += 1
            if bad <= 5:
                print("   MISMATCH", i, rays[i], "mine", s[i], m[i], res["rot"][i], "ref", rs, rm, rl)
        else:
            maxrel = max(maxrel, abs(res["rot"][i] - rl) / max(1e-30, abs(rl)))
    print(f"  surface/modifier mismatches: {bad} of {len(ref)}; max rel L diff (6 digits printed) {maxrel:.2e}")


def check_bins():
    g = json.load(open(G / "golden.json"))
    up = np.load(G / "bin_dirs.npy")
    for name in [k for k in g if k.startswith("bins_")]:
        args = g[name]["args"]
        ctx = _lib.Context(0, _lib.RB_PROGRAM_RCONTRIB)
        ctx.load_octree(G / "contrib.oct")
        ctx.set_options(["-ab", "0"])
        i = 0
        params, binv, bn = "", "0", "1"
        while i < len(args):
            a = args[i]
            if a == "-f": ctx.cal_load(args[i + 1])
            elif a == "-e": ctx.cal_set(args[i + 1])
            elif a == "-p": params = args[i + 1]; ctx.cal_set(params)
            elif a == "-bn": bn = args[i + 1]
            elif a == "-b": binv = args[i + 1]
            elif a == "-m": ctx.add_modifier(args[i + 1], params, binv, int(ctx.cal_eval(bn) + .5))
            i += 2
        m = ctx.rcontrib(up)
        bins = np.where(m[:, :, 0].sum(axis=1) > 0, m[:, :, 0].argmax(axis=1), -1)
        ref = np.array(g[name]["bins"])
        print(f"  {name}: cols {m.shape[1]} bin mismatches {(bins != ref).sum()} rowsum minmax {m[:, :, 0].sum(1).min()} {m[:, :, 0].sum(1).max()}")


RB_ARGS = ["-f", "reinhartb.cal", "-p", "MF=1,rNx=0,rNy=0,rNz=-1,Ux=0,Uy=1,Uz=0,RHS=+1", "-bn", "Nrbins", "-b", "rbin",
           "-m", "skyglow"]


def setup_rc(octf, opts, mf=1):
    ctx = _lib.Context(0, _lib.RB_PROGRAM_RCONTRIB)
    ctx.load_octree(octf)
    ctx.set_options(opts)
    ctx.cal_load("reinhartb.cal")
    p = f"MF={mf},rNx=0,rNy=0,rNz=-1,Ux=0,Uy=1,Uz=0,RHS=+1"
    ctx.cal_set(p)
    ctx.add_modifier("skyglow", p, "rbin", int(ctx.cal_eval("Nrbins") + .5))
    return ctx


def check_rcontrib_small():
    sens = np.array([[10, 10, 3, 0, 0, 1], [4, 5, 3, 0, 0, 1], [20, 20, 12, 0, 0, 1]], dtype=float)
    opts = ["-ab", "1", "-ad", "4096", "-lw", "1e-4"]
    ctx = setup_rc(G / "contrib.oct", opts)
    m = ctx.rcontrib(sens, flags=_lib.RB_IRRAD_RCONTRIB)
    ref = refrun.rcontrib(G / "contrib.oct", sens, ["-I"] + opts + RB_ARGS).reshape(3, -1, 3)
    print("  contrib.oct -ab 1 row sums mine", m[:, :, 0].sum(1), "ref", ref[:, :, 0].sum(1), "stats", ctx.stats())


def check_rcontrib_office(npolys=100000, nsens=512, ab=3, ad=4096):
    rad = TMP / f"off{npolys}.rad"
    octf = TMP / f"off{npolys}.oct"
    t = time.time()
    scenegen.write_office(rad, npolys=npolys, seed=1234)
    t1 = time.time()
    scenegen.build_octree(rad, octf)
    print(f"office{npolys}: gen {t1 - t:.2f}s oconv {time.time() - t1:.2f}s size {octf.stat().st_size / 1e6:.1f} MB")
    sens = scenegen.office_sensors(nsens)
    opts = ["-ab", str(ab), "-ad", str(ad), "-lw", f"{1.0 / ad:.3e}"]
    ctx = setup_rc(octf, opts)
    t = time.time()
    m = ctx.rcontrib(sens, flags=_lib.RB_IRRAD_RCONTRIB)
    dt = time.time() - t
    st = ctx.stats()
    print(f"  mine: {nsens} sensors in {dt:.3f}s; rays {st['nrays']} -> {st['nrays'] / dt / 1e6:.1f} Mrays/s wall, "
          f"{st['nrays'] / (st['wave_ms'] / 1e3) / 1e6:.1f} Mrays/s in k_wave; stats {st}")
    ctx.reset_stats()
    t = time.time()
    m2 = ctx.rcontrib(sens, flags=_lib.RB_IRRAD_RCONTRIB)
    dt = time.time() - t
    st = ctx.stats()
    print(f"  mine (2nd): {dt:.3f}s; {st['nrays'] / dt / 1e6:.1f} Mrays/s wall; k_wave {st['wave_ms']:.1f} ms of kernel {st['kernel_ms']:.1f} ms")
    nref = min(nsens, 64)
    t = time.time()
    ref = refrun.rcontrib(octf, sens[:nref], ["-I"] + opts + RB_ARGS, nproc=os.cpu_count()).reshape(nref, -1, 3)
    dtr = time.time() - t
    print(f"  ref : {nref} sensors in {dtr:.2f}s with {os.cpu_count()} procs")
    a = m[:nref, :, 0].sum(1)
    b = ref[:, :, 0].sum(1)
    print("  row sums mine", a[:8], "\n  row sums ref ", b[:8])
    print("  mean row sum mine %.5f ref %.5f ; total-matrix rel diff %.4f" % (a.mean(), b.mean(), abs(a.sum() - b.sum()) / b.sum()))
    # per-bin comparison pooled over sensors
    pa = m[:nref, :, 0].sum(0)
    pb = ref[:, :, 0].sum(0)
    nz = pb > 0
    print("  pooled per-bin rel diff: median %.4f max %.4f" % (np.median(np.abs(pa[nz] - pb[nz]) / pb[nz]),
                                                             np.max(np.abs(pa[nz] - pb[nz]) / pb[nz])))


if __name__ == "__main__":
    quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
    steps = [check_known, check_bins, check_office, check_rcontrib_small]
    if not quick:
        steps += [lambda: check_office(100000, 500000), check_rcontrib_office]
    for f in steps:
        try:
            f()
        except Exception:
            traceback.print_exc()
