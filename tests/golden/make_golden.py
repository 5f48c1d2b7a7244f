"""Regenerates the golden fixtures of tests/golden/ (run in the build container,
where /root/reference and oracle/_ref exist):

  * trace.oct, contrib.oct  -- the reference's own frozen test octrees
    (/root/reference/tests/Resources), input fixtures only
  * *.json                  -- known-answer outputs of the UNMODIFIED reference
    rtrace / rcontrib (oracle/_ref/bin) on seeded inputs

Usage: python tests/golden/make_golden.py
"""
import json
import shutil
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path.insert(0, str(ROOT))
from oracle import refrun  # noqa: E402

RES = Path("/root/reference/tests/Resources")


def main():
    for f in ("trace.oct", "contrib.oct"):
        shutil.copyfile(RES / f, HERE / f)
    g = {}
    # SURVEY 8(c) known-answer vectors
    rays = np.array([[1, 2, 3, 0, 0, 1], [4, 5, 6, 0, 1, 0], [10, 10, 3, 0, 0, -1]], dtype=float)
    g["trace_ovposmNL"] = {"rays": rays.tolist(),
                           "args": ["-ab", "0", "-ovposmNL"],
                           "out": refrun.rtrace(HERE / "trace.oct", rays, ["-ab", "0", "-ovposmNL"])}
    sens = np.array([[10, 10, .5, 0, 0, 1], [20, 20, .5, 0, 0, 1], [20, 20, 9.5, 0, 0, 1],
                     [1, 2, 2.5, 0, 0, 1], [39, 45, 9.1, 0, 0, 1], [5, 40, 12, 0.3, 0.1, 0.9]], dtype=float)
    g["trace_I_ab0"] = {"rays": sens.tolist(), "args": ["-I", "-ab", "0"],
                        "out": refrun.rtrace(HERE / "trace.oct", sens, ["-I", "-ab", "0"])}
    # 100x100 sensor grid of config 1 (S-small): values as float64
    gx, gy = np.meshgrid(np.linspace(1, 39, 100), np.linspace(2, 45, 100))
    grid = np.stack([gx.ravel(), gy.ravel(), np.full(10000, 2.5), np.zeros(10000), np.zeros(10000),
                     np.ones(10000)], axis=1)
    v = refrun.rtrace(HERE / "trace.oct", grid, ["-I", "-ab", "0", "-dt", "0", "-dj", "0", "-dc", "1"], outform="d")
    v = v.reshape(-1, 3)
    g["trace_grid_I_ab0"] = {"nonzero_rows": int((v[:, 0] > 0).sum()), "sum": float(v.sum()),
                             "unique": sorted(set(np.round(v[:, 0], 6).tolist()))[:8]}
    # deterministic rcontrib: -ab 0 from above the scene, one ray -> one bin (SURVEY 8a a19 check)
    rng = np.random.default_rng(5)
    d = rng.normal(size=(3000, 3))
    d[:, 2] = np.abs(d[:, 2]) + 0.01
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    up = np.concatenate([np.tile([20., 20., 20.], (3000, 1)), d], axis=1)
    cases = {
        "reinhartb_mf1": ["-f", "reinhartb.cal", "-p", "MF=1,rNx=0,rNy=0,rNz=-1,Ux=0,Uy=1,Uz=0,RHS=+1", "-bn", "Nrbins", "-b", "rbin", "-m", "skyglow"],
        "reinhartb_mf4": ["-f", "reinhartb.cal", "-p", "MF=4,rNx=0,rNy=0,rNz=-1,Ux=0,Uy=1,Uz=0,RHS=+1", "-bn", "Nrbins", "-b", "rbin", "-m", "skyglow"],
        "reinhart_mf2": ["-e", "MF:2", "-f", "reinhart.cal", "-b", "rbin", "-bn", "Nrbins", "-m", "skyglow"],
        "klems_full": ["-f", "klems_full.cal", "-p", "RHS=+1", "-bn", "Nkbins", "-b", "kbin(0,0,-1,0,1,0)", "-m", "skyglow"],
        "klems_half": ["-f", "klems_half.cal", "-bn", "Nkhbins", "-b", "khbin(0,0,-1,0,1,0)", "-m", "skyglow"],
        "klems_quarter": ["-f", "klems_quarter.cal", "-bn", "Nkqbins", "-b", "kqbin(0,0,-1,0,1,0)", "-m", "skyglow"],
        "hemi": ["-b", "if(-Dx*0-Dy*0-Dz*-1,0,-1)", "-bn", "1", "-m", "skyglow"],
        "shirchiu": ["-f", "disk2square.cal", "-p", "SCdim=6,rNx=0,rNy=0,rNz=-1,Ux=0,Uy=1,Uz=0,RHS=+1", "-bn", "SCdim*SCdim",
                     "-b", "scbin", "-m", "skyglow"],
    }
    np.save(HERE / "bin_dirs.npy", up)
    for name, args in cases.items():
        m = refrun.rcontrib(HERE / "contrib.oct", up, ["-ab", "0"] + args)
        m = m.reshape(3000, -1, 3)
        bins = np.where(m[:, :, 0].sum(axis=1) > 0, m[:, :, 0].argmax(axis=1), -1)
        g["bins_" + name] = {"args": args, "ncols": int(m.shape[1]), "bins": bins.astype(int).tolist()}
    # instances and meshes: fixtures built by the reference oconv / obj2mesh from
    # the small text scenes in tests/golden/volumes/, known answers by the reference rtrace
    import os
    import subprocess
    V = HERE / "volumes"
    env = dict(os.environ, RAYPATH=f".:{refrun.LIB}")
    def sh(cmd, out=None):
        r = subprocess.run(cmd, cwd=V, env=env, capture_output=True)
        assert r.returncode == 0, r.stderr
        if out:
            (V / out).write_bytes(r.stdout)
    sh([str(refrun.BIN / "oconv"), "-f", "louvre.rad"], "louvre.oct")
    sh([str(refrun.BIN / "obj2mesh"), "-a", "meshmats.rad", "bump.obj", "bump.rtm"])
    sh([str(refrun.BIN / "oconv"), "-f", "room.rad"], "room.oct")
    sh([str(refrun.BIN / "oconv"), "-f", "meshroom.rad"], "meshroom.oct")
    rng = np.random.default_rng(6)
    o = rng.uniform((-4.5, -4.5, -0.9), (4.5, 4.5, 3.5), size=(4000, 3))
    d = rng.normal(size=(4000, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    vr = np.concatenate([o, d], axis=1)
    np.save(HERE / "volume_rays.npy", vr)
    for name in ("room", "meshroom"):
        r = subprocess.run([str(refrun.BIN / "rtrace"), "-h", "-fda", "-ab", "0", "-osmLN", f"{name}.oct"], cwd=V, env=env,
                           input=vr.tobytes(), capture_output=True)
        assert r.returncode == 0, r.stderr
        g["volumes_" + name] = r.stdout.decode()
    # 5-phase direct-sun matrix in miniature (BASELINE config 5): 145 `light` suns at the Reinhart MF:1
    # patch centres sharing modifier `solar`, louvre instances + meshes shading a floor of sensors
    import io
    from pyradiance_b200 import scenegen
    buf = io.StringIO()
    buf.write((V / "sunroom_base.rad").read_text())
    scenegen.write_suns(buf, mf=1)
    (V / "sunroom.rad").write_text(buf.getvalue())
    sh([str(refrun.BIN / "oconv"), "-f", "sunroom.rad"], "sunroom.oct")
    sx, sy = np.meshgrid(np.linspace(-4, 4, 8), np.linspace(-4, 4, 6))
    ss = np.stack([sx.ravel(), sy.ravel(), np.full(48, 0.01), np.zeros(48), np.zeros(48), np.ones(48)], axis=1)
    np.save(HERE / "sunroom_sensors.npy", ss)
    sun_args = ["-I+", "-ab", "0", "-dc", "1", "-dt", "0", "-dj", "0", "-e", "MF:1", "-f", "reinhart.cal",
                "-b", "rbin", "-bn", "Nrbins", "-m", "solar"]
    r = subprocess.run([str(refrun.BIN / "rcontrib"), "-h", "-fdd"] + sun_args + ["sunroom.oct"], cwd=V, env=env,
                       input=ss.tobytes(), capture_output=True)
    assert r.returncode == 0, r.stderr
    np.save(HERE / "sunroom_ab0.npy", np.frombuffer(r.stdout, dtype=np.float64).reshape(48, 146, 3))
    g["sunroom_args"] = sun_args
    # local light sources: polygon / triangle / pentagon / sphere / ring / cylinder
    # emitters, an illum, a glow with a radius, a spotlight, occluders and a glass screen.  Deterministic
    # reference runs (-u- -dj 0 -dt 0 -dc 1): irradiance at 500 sensors for three -ds settings, values seen
    # by 500 view rays, and rcontrib coefficients per source modifier.
    L = HERE / "lights"
    r = subprocess.run([str(refrun.BIN / "oconv"), "-f", "lights.rad"], cwd=L, env=env, capture_output=True)
    assert r.returncode == 0, r.stderr
    (L / "lights.oct").write_bytes(r.stdout)
    rng = np.random.default_rng(3)
    n = 400
    pts = np.stack([rng.uniform(0.2, 7.8, n), rng.uniform(0.2, 5.8, n), np.full(n, 0.02)], 1)
    sens = np.concatenate([pts, np.tile([0, 0, 1.], (n, 1))], 1)
    pw = np.stack([np.full(100, 7.98), rng.uniform(0.2, 5.8, 100), rng.uniform(0.2, 2.8, 100)], 1)
    sens = np.concatenate([sens, np.concatenate([pw, np.tile([-1., 0, 0], (100, 1))], 1)])
    o = rng.uniform((0.5, 0.5, 0.3), (7.5, 5.5, 2.5), size=(500, 3))
    d = rng.normal(size=(500, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays = np.concatenate([o, d], 1)
    out = {"sensors": sens, "rays": rays}
    det = ["-u-", "-ab", "0", "-dt", "0", "-dj", "0", "-dc", "1"]
    for ds in ("0.2", "0", "0.05"):
        out["irrad_ds" + ds] = refrun.rtrace(L / "lights.oct", sens, ["-I"] + det + ["-ds", ds], outform="d").reshape(-1, 3)
    out["view_ds0.2"] = refrun.rtrace(L / "lights.oct", rays, det + ["-ds", ".2"], outform="d").reshape(-1, 3)
    mods = ["lum", "lum2", "spot", "glw", "ill"]
    margs = ["-u-", "-I", "-ab", "0", "-dj", "0", "-ds", ".2"]
    for m in mods:
        margs += ["-m", m]
    out["rcontrib_ab0"] = refrun.rcontrib(L / "lights.oct", sens, margs).reshape(len(sens), -1, 3)
    # the same two runs in the 4-byte RGBE output format (-f?c)
    r = subprocess.run([str(refrun.BIN / "rtrace"), "-h", "-fdc", "-ov"] + det + ["-ds", ".2", "lights.oct"], cwd=L, env=env,
                       input=rays.tobytes(), capture_output=True)
    assert r.returncode == 0, r.stderr
    out["view_rgbe"] = np.frombuffer(r.stdout, dtype=np.uint8).reshape(-1, 4)
    r = subprocess.run([str(refrun.BIN / "rcontrib"), "-h", "-fdc"] + margs + ["lights.oct"], cwd=L, env=env,
                       input=sens.tobytes(), capture_output=True)
    assert r.returncode == 0, r.stderr
    out["rcontrib_rgbe"] = np.frombuffer(r.stdout, dtype=np.uint8).reshape(-1, 4)
    np.savez_compressed(HERE / "lights.npz", **out)
    g["lights_mods"] = mods
    # sky brightness patterns under glow emitters (brightfunc skybright.cal `skybr`, perezlum.cal
    # `skybright`): values seen by 2000 rays in all directions, for the reference's own
    # trace.oct (perezlum), a rotated sunny sky and an overcast / intermediate pair
    S = HERE / "sky"
    rng = np.random.default_rng(1)
    d = rng.normal(size=(2000, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    skyrays = np.concatenate([np.tile([20., 20., 20.], (2000, 1)), d], 1)
    sky = {"rays": skyrays}
    sky["trace"] = refrun.rtrace(HERE / "trace.oct", skyrays, ["-ab", "0"], outform="d").reshape(-1, 3)
    for name in ("skies", "overcast"):
        r = subprocess.run([str(refrun.BIN / "oconv"), "-f", f"{name}.rad"], cwd=S, env=env, capture_output=True)
        assert r.returncode == 0, r.stderr
        (S / f"{name}.oct").write_bytes(r.stdout)
        sky[name] = refrun.rtrace(S / f"{name}.oct", skyrays, ["-ab", "0"], outform="d").reshape(-1, 3)
    np.savez_compressed(HERE / "sky.npz", **sky)
    # rfluxmtx front-end: the rcontrib command lines the reference synthesises (-v), a
    # deterministic pass-through matrix (view rays -> Klems window) and a sender-sampled daylight matrix
    # (Klems window -> Reinhart MF:2 sky, 5000 samples per sender bin) as the statistical truth
    import shlex
    F = HERE / "flux"
    fenv = dict(env, PATH=f"{refrun.BIN}:{os.environ['PATH']}")
    specs = ["- window_kf.rad room.rad", "-ab 1 -ad 64 -lw 1e-2 -I+ -y 3 - window_multi.rad room.rad",
             "-w -ab 2 -c 4 - sky_r2.rad room.rad window_kf.rad", "-ffd -ab 0 -c 50 sender_window.rad sky_r2.rad room.rad",
             "-bj .7 -fdf -c 3 -ab 1 - window_kf.rad -i x.oct room.rad"]
    cmds = []
    for sp in specs:
        r = subprocess.run([str(refrun.BIN / "rfluxmtx"), "-v"] + sp.split(), cwd=F, env=fenv, input=b"", capture_output=True)
        line = [ln for ln in r.stderr.decode().splitlines() if "rcontrib -fo+" in ln][0]
        cmds.append({"spec": sp, "rcontrib": shlex.split(line[line.index("rcontrib -fo+"):])[:-1]})
    g["rfluxmtx_commands"] = cmds
    rng = np.random.default_rng(8)
    o = rng.uniform((0.5, 1.0, 0.5), (3.5, 4.5, 2.5), size=(300, 3))
    d = rng.normal(size=(300, 3))
    d[:, 1] = -np.abs(d[:, 1]) - 0.5                       # towards the south wall with the window
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    frays = np.concatenate([o, d], 1)
    flux = {"rays": frays}
    r = subprocess.run([str(refrun.BIN / "rfluxmtx"), "-h", "-fdd", "-ab", "0", "-", "window_kf.rad", "room.rad"], cwd=F,
                       env=fenv, input=frays.tobytes(), capture_output=True)
    assert r.returncode == 0, r.stderr
    flux["pass_kf"] = np.frombuffer(r.stdout, dtype=np.float64).reshape(300, 145, 3)
    r = subprocess.run([str(refrun.BIN / "rfluxmtx"), "-h", "-ffd", "-ab", "0", "-c", "5000", "sender_window.rad", "sky_r2.rad",
                        "room.rad"], cwd=F, env=fenv, capture_output=True)
    assert r.returncode == 0, r.stderr
    flux["dmx_kf_r2"] = np.frombuffer(r.stdout, dtype=np.float64).reshape(145, 578, 3)[:, :, 0].astype(np.float32)
    np.savez_compressed(HERE / "flux.npz", **flux)
    # vwrays: rays of every view type from the reference binary, -fd, plus -d dimensions
    vcases = {
        "persp": ["-vtv", "-vp", "2", "3", "1.5", "-vd", "0.3", "1", "-0.1", "-vu", "0", "0", "1", "-vh", "60", "-vv", "40"],
        "persp_shift_clip": ["-vtv", "-vp", "2", "3", "1.5", "-vd", "1", "0.2", "0", "-vu", "0", "0", "1", "-vh", "50", "-vv", "50",
                             "-vs", "0.2", "-vl", "-0.1", "-vo", "0.5", "-va", "10"],
        "parallel": ["-vtl", "-vp", "0", "0", "10", "-vd", "0", "0", "-1", "-vu", "0", "1", "0", "-vh", "8", "-vv", "6"],
        "fisheye_h": ["-vth", "-vp", "2", "2", "1", "-vd", "0", "-1", "0", "-vu", "0", "0", "1", "-vh", "180", "-vv", "180"],
        "fisheye_a": ["-vta", "-vp", "2", "2", "1", "-vd", "0", "0", "1", "-vu", "0", "1", "0", "-vh", "180", "-vv", "180"],
        "cyl": ["-vtc", "-vp", "2", "2", "1", "-vd", "1", "0", "0", "-vu", "0", "0", "1", "-vh", "300", "-vv", "60"],
        "plan": ["-vts", "-vp", "2", "2", "1", "-vd", "0", "0", "1", "-vu", "0", "1", "0", "-vh", "200", "-vv", "200"],
    }
    vw = {}
    g["vwrays"] = {}
    for name, va in vcases.items():
        vw[name] = np.frombuffer(refrun.run("vwrays", ["-fd", "-x", "24", "-y", "18"] + va), dtype=np.float64).reshape(-1, 6)
        g["vwrays"][name] = {"view": va, "dim": refrun.run("vwrays", ["-d", "-x", "24", "-y", "18"] + va).decode(),
                             "ascii_5x4": refrun.run("vwrays", ["-x", "5", "-y", "4"] + va).decode()}
    np.savez_compressed(HERE / "vwrays.npz", **vw)
    # dctimestep: small matrix files (ascii DC with header, float sky with header, headerless
    # ascii sky for -n) and the reference's products, ascii text and float bytes; plus a V.T.D.s chain
    D = HERE / "dct"
    D.mkdir(exist_ok=True)
    rng = np.random.default_rng(12)

    def write_mtx(path, m, fmt):
        hdr = f"#?RADIANCE\nNROWS={m.shape[0]}\nNCOLS={m.shape[1]}\nNCOMP=3\n"
        if fmt == "a":
            body = "".join("\t".join("%.6e %.6e %.6e" % tuple(c) for c in row) + "\n" for row in m).encode()
            path.write_bytes((hdr + "FORMAT=ascii\n\n").encode() + body)
        else:
            dt = np.float32 if fmt == "f" else np.float64
            path.write_bytes((hdr + "BigEndian=0\nFORMAT=" + ("float" if fmt == "f" else "double") + "\n\n").encode()
                             + np.ascontiguousarray(m, dtype=dt).tobytes())
    dc = (rng.random((37, 146, 3)) ** 6 * 0.05).astype(np.float32)          # coefficients over several decades
    sky = (rng.random((146, 29, 3)) ** 3 * 2e4).astype(np.float32)
    sky[:, 5] = 0                                                           # a night-time step
    write_mtx(D / "dc.mtx", dc, "a")
    write_mtx(D / "sky_f.smx", sky, "f")
    write_mtx(D / "sky_d.smx", sky, "d")
    (D / "sky_n.txt").write_text("".join("%.6e %.6e %.6e\n" % tuple(c) for row in sky for c in row))
    vm = (rng.random((11, 20, 3)) * 0.2).astype(np.float32)
    tm = (rng.random((20, 20, 3)) * 0.1).astype(np.float32)
    dm = (rng.random((20, 146, 3)) * 0.01).astype(np.float32)
    write_mtx(D / "v.mtx", vm, "f"); write_mtx(D / "t.mtx", tm, "a"); write_mtx(D / "d.mtx", dm, "d")
    dct = {}

    def body(b):
        return b[b.index(b"\n\n") + 2:]
    r = subprocess.run([str(refrun.BIN / "dctimestep"), "dc.mtx", "sky_f.smx"], cwd=D, env=env, capture_output=True)
    assert r.returncode == 0, r.stderr
    g["dctimestep_ascii"] = body(r.stdout).decode()
    g["dctimestep_header"] = r.stdout[:r.stdout.index(b"\n\n")].decode()
    r = subprocess.run([str(refrun.BIN / "dctimestep"), "-of", "dc.mtx", "sky_d.smx"], cwd=D, env=env, capture_output=True)
    dct["dc_sky"] = np.frombuffer(body(r.stdout), dtype=np.float32).reshape(37, 29, 3)
    r = subprocess.run([str(refrun.BIN / "dctimestep"), "-h", "-of", "-n", "29", "dc.mtx", "sky_n.txt"], cwd=D, env=env, capture_output=True)
    assert r.returncode == 0, r.stderr
    dct["dc_sky_n"] = np.frombuffer(r.stdout, dtype=np.float32).reshape(37, 29, 3)
    r = subprocess.run([str(refrun.BIN / "dctimestep"), "-h", "-of", "v.mtx", "t.mtx", "d.mtx", "sky_f.smx"], cwd=D, env=env, capture_output=True)
    assert r.returncode == 0, r.stderr
    dct["vtds"] = np.frombuffer(r.stdout, dtype=np.float32).reshape(11, 29, 3)
    np.savez_compressed(HERE / "dct.npz", **dct)
    (HERE / "golden.json").write_text(json.dumps(g, indent=0))
    print("wrote", HERE / "golden.json")


if __name__ == "__main__":
    main()
