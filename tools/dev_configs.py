"""Developer check of BASELINE configs 3 and 4 at reduced size against the reference binaries."""
import os, sys, time
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from pyradiance_b200 import _lib, scenegen
from oracle import refrun
TMP = Path(os.environ.get("RB_TMP", "/tmp/rbt")); TMP.mkdir(parents=True, exist_ok=True)

def c3(npoly=1_000_000, floors=10, nsens=2000, nref=24):
    rad, octf = TMP / "bld.rad", TMP / "bld.oct"
    t = time.time(); scenegen.write_office(rad, npolys=npoly, floors=floors, seed=77); t1 = time.time()
    scenegen.build_octree(rad, octf); print(f"C3 scene: gen {t1-t:.1f}s oconv {time.time()-t1:.1f}s {octf.stat().st_size/1e6:.0f} MB")
    sens = scenegen.office_sensors(nsens, floors=floors, seed=3)
    opts = ["-ab", "5", "-ad", "10000", "-lw", "1e-4"]
    P = "MF=4,rNx=0,rNy=0,rNz=-1,Ux=0,Uy=1,Uz=0,RHS=+1"
    ctx = _lib.Context(0, _lib.RB_PROGRAM_RCONTRIB); t = time.time(); ctx.load_octree(octf); print(f"  load {time.time()-t:.2f}s")
    ctx.set_options(opts); ctx.cal_load("reinhartb.cal"); ctx.cal_set(P); ctx.add_modifier("skyglow", P, "rbin", int(ctx.cal_eval("Nrbins")+.5))
    t = time.time(); m = ctx.rcontrib(sens, flags=_lib.RB_IRRAD_RCONTRIB); dt = time.time()-t; st = ctx.stats()
    print(f"  GPU: {nsens} sensors x 2305 bins in {dt:.2f}s, {st['nrays']/1e6:.0f} Mrays, {st['nrays']/dt/1e6:.0f} Mrays/s wall, k_trace {st['nrays']/st['wave_ms']/1e3:.0f} Mrays/s, batches {st['batches']} retries {st['retries']}; nodes/ray {st['nodes']/st['nrays']:.1f} prims/ray {st['prims']/st['nrays']:.1f}")
    idx = np.linspace(0, nsens-1, nref).astype(int)
    t = time.time()
    ref = refrun.rcontrib(octf, sens[idx], ["-I+"] + opts + ["-f", "reinhartb.cal", "-p", P, "-bn", "Nrbins", "-b", "rbin", "-m", "skyglow"], nproc=os.cpu_count()).reshape(nref, -1, 3)
    print(f"  ref: {nref} sensors in {time.time()-t:.1f}s")
    a = m[idx, :, 0].sum(1); b = ref[:, :, 0].sum(1)
    print("  row sums gpu", np.round(a[:8], 4), "\n  row sums ref", np.round(b[:8], 4), "\n  total gpu %.4f ref %.4f" % (a.sum(), b.sum()))

def c4(nrays=200000, nref=2000):
    # S-view-like: 8 window groups, each window its own glow modifier, Klems full bins
    rad, octf = TMP / "view.rad", TMP / "view.oct"
    import io
    rng = np.random.default_rng(5); out = io.StringIO(); out.write(scenegen.MATERIALS)
    for i in range(8):
        out.write(f"void glow wg{i}\n0\n0\n4 1 1 1 0\n\n")
    # office with windows replaced by glow polygons wg_i (facing inward)
    buf = io.StringIO(); scenegen.office_floor(buf, rng, 0.0, 1500, tag="f0")
    txt = buf.getvalue()
    for i in range(8):
        txt = txt.replace(f"win_glass polygon f0.win{i}\n", f"wg{i} polygon f0.win{i}\n")
    out.write(txt); rad.write_text(out.getvalue()); scenegen.build_octree(rad, octf)
    rays = scenegen.random_rays(nrays, seed=9, lo=(5, 5, 0.5), hi=(35, 20, 2.5))
    opts = ["-ab", "3", "-ad", "1024", "-lw", "1e-4"]
    ctx = _lib.Context(0, _lib.RB_PROGRAM_RCONTRIB); ctx.load_octree(octf); ctx.set_options(opts)
    ctx.cal_load("klems_full.cal"); ctx.cal_set("RHS=+1")
    args = ["-f", "klems_full.cal", "-p", "RHS=+1", "-bn", "Nkbins", "-b", "kbin(0,-1,0,0,0,1)"]
    for i in range(8):
        ctx.add_modifier(f"wg{i}", "RHS=+1", "kbin(0,-1,0,0,0,1)", int(ctx.cal_eval("Nkbins")+.5)); args += ["-m", f"wg{i}"]
    t = time.time(); m = ctx.rcontrib(rays); dt = time.time()-t; st = ctx.stats()
    print(f"C4-mini GPU: {nrays} view rays x {ctx.num_columns()} cols in {dt:.2f}s, {st['nrays']/1e6:.0f} Mrays, {st['nrays']/dt/1e6:.0f} Mrays/s wall")
    t = time.time(); ref = refrun.rcontrib(octf, rays[:nref], opts + args, nproc=os.cpu_count()).reshape(nref, -1, 3); print(f"  ref {nref} rays {time.time()-t:.1f}s")
    g = m[:nref, :, 0].astype(float); r = ref[:, :, 0]
    per_mod_g = g.reshape(nref, 8, 145).sum((0, 2)); per_mod_r = r.reshape(nref, 8, 145).sum((0, 2))
    print("  per-window-group totals gpu", np.round(per_mod_g, 2), "\n  per-window-group totals ref", np.round(per_mod_r, 2))
    print("  total gpu %.3f ref %.3f" % (g.sum(), r.sum()))

def c5(npoly=100_000, nsens=int(os.environ.get("C5_SENSORS", 1000)), nref=8, mf=6):
    # S-sun: MF:6 light suns sharing modifier solar, -ab 1, facade louvres as octree instances + two meshes
    # (needs the reference oconv for the instances; without it the plain scene is used)
    import io, shutil
    rad, octf = TMP / "sun.rad", TMP / "sun.oct"
    out = io.StringIO(); out.write(scenegen.MATERIALS); scenegen.write_suns(out, mf=mf)
    rng = np.random.default_rng(11); scenegen.office_floor(out, rng, 0.0, (npoly - 24) // 6, tag="f0")
    vol = ROOT / "tests" / "golden" / "volumes"
    os.environ["RB_RAYPATH_EXTRA"] = str(TMP)
    if refrun.available():
        for f in ("louvre.oct", "bump.rtm"):
            shutil.copyfile(vol / f, TMP / f)
        for i in range(64):              # louvres outside the south windows, 8 per window
            x, z = 1.0 + 38.0 * (i + 0.5) / 64, 1.0 + 0.25 * (i % 8)
            out.write(f"void instance lv{i}\n7 louvre.oct -s 0.6 -t {x:.3f} -0.45 {z:.3f}\n0\n0\n\n")
        out.write("void mesh bumpA\n7 bump.rtm -s 1.5 -t 10 12 0.75\n0\n0\n\nvoid mesh bumpB\n9 bump.rtm -rz 40 -s 2 -t 28 9 0.75\n0\n0\n\n")
        rad.write_text(out.getvalue())
        scenegen.build_octree(rad, octf, use_reference_oconv=str(refrun.BIN / "oconv"))
    else:
        rad.write_text(out.getvalue()); scenegen.build_octree(rad, octf)
    sens = scenegen.office_sensors(nsens, seed=4)
    opts = ["-ab", "1", "-ad", "256", "-lw", "1e-3", "-dc", "1", "-dt", "0", "-dj", "0"]
    nb = 144 * mf * mf + 2
    ctx = _lib.Context(0, _lib.RB_PROGRAM_RCONTRIB); ctx.load_octree(octf); ctx.set_options(opts)
    ctx.cal_load("reinhart.cal"); ctx.cal_set(f"MF={mf}"); ctx.add_modifier("solar", "", "rbin", nb)
    for rep in range(2):
        ctx.reset_stats(); t = time.time(); m = ctx.rcontrib(sens, flags=_lib.RB_IRRAD_RCONTRIB); dt = time.time() - t; st = ctx.stats()
        print(f"C5-mini GPU: {nsens} sensors x {nb} bins in {dt:.2f}s, {st['nrays']/1e6:.0f} Mrays, {st['nrays']/dt/1e6:.0f} Mrays/s wall, k_trace {st['nrays']/max(st['wave_ms'],1e-9)/1e3:.0f} Mrays/s, shade_ms {st['shade_ms']:.0f} wave_ms {st['wave_ms']:.0f} batches {st['batches']} retries {st['retries']} waves {st['waves']}")
    idx = np.linspace(0, nsens - 1, nref).astype(int)
    t = time.time()
    ref = refrun.rcontrib(octf, sens[idx], ["-I+"] + opts + ["-e", f"MF:{mf}", "-f", "reinhart.cal", "-b", "rbin", "-bn", "Nrbins", "-m", "solar"], nproc=os.cpu_count()).reshape(nref, -1, 3)
    print(f"  ref: {nref} sensors in {time.time()-t:.1f}s")
    a = m[idx, :, 0].sum(1); b = ref[:, :, 0].sum(1)
    print("  row sums gpu", np.round(a, 5), "\n  row sums ref", np.round(b, 5), "\n  total gpu %.5f ref %.5f" % (a.sum(), b.sum()))


if __name__ == "__main__":
    which = sys.argv[1:] or ["c3", "c4"]
    if "c4" in which: c4()
    if "c3" in which: c3()
    if "c5" in which: c5()
