"""Parity tests proper: the CUDA path, called through the C ABI (ctypes), against
the CPU oracle (oracle/rb_oracle.c), the golden vectors made by the unmodified
reference, and -- when oracle/_ref travelled to this box -- the reference
rtrace / rcontrib binaries themselves, on the same seeded inputs.

Bars (BASELINE.json north_star): hit surface and modifier bit-exact; -ab 0
values within 1e-5 relative; stochastic -ab > 0 coefficients within the
statistical tolerance written in each test."""
import json

import numpy as np
import pytest

from oracle import port, refrun
from pyradiance_b200 import _lib, scenegen
import pyradiance_b200 as pr

pytestmark = pytest.mark.gpu

RB_P = "MF=1,rNx=0,rNy=0,rNz=-1,Ux=0,Uy=1,Uz=0,RHS=+1"
RB_ARGS = ["-f", "reinhartb.cal", "-p", RB_P, "-bn", "Nrbins", "-b", "rbin", "-m", "skyglow"]


@pytest.fixture(scope="module")
def G(golden):
    return json.load(open(golden / "golden.json"))


@pytest.fixture(scope="module")
def office2k(workdir):
    rad, octf = workdir / "off2k.rad", workdir / "off2k.oct"
    scenegen.write_office(rad, npolys=2000, seed=1)
    scenegen.build_octree(rad, octf)
    return octf


@pytest.fixture(scope="module")
def office100k(workdir):
    rad, octf = workdir / "off100k.rad", workdir / "off100k.oct"
    scenegen.write_office(rad, npolys=100_000, seed=1234)
    scenegen.build_octree(rad, octf)
    return octf


def names(ctx, idx):
    return [ctx.object_name(int(i)) if i >= 0 else "*" for i in idx]


def rc_ctx(octf, opts, mf=1, mod="skyglow"):
    ctx = _lib.Context(0, _lib.RB_PROGRAM_RCONTRIB)
    ctx.load_octree(octf)
    ctx.set_options(opts)
    ctx.cal_load("reinhartb.cal")
    p = RB_P.replace("MF=1", f"MF={mf}")
    ctx.cal_set(p)
    ctx.add_modifier(mod, p, "rbin", int(ctx.cal_eval("Nrbins") + .5))
    return ctx


# ------------------------------------------------------------ deterministic --
def test_known_answer_vectors(G, golden):
    k = G["trace_ovposmNL"]
    ctx = _lib.Context(0)
    ctx.load_octree(golden / "trace.oct")
    ctx.set_options(["-ab", "0"])
    _, res = ctx.rtrace(np.array(k["rays"]), want_values=False)
    s, m = names(ctx, res["robj"]), names(ctx, res["omod"])
    for i, line in enumerate(k["out"].strip("\n").split("\n")):
        f = line.split("\t")
        assert (s[i], m[i]) == (f[9], f[10])
        np.testing.assert_allclose(res["rop"][i], [float(x) for x in f[3:6]], rtol=2e-7, atol=1e-12)
        np.testing.assert_allclose(res["ron"][i], [float(x) for x in f[11:14]], atol=1e-9)
        assert res["rot"][i] == pytest.approx(float(f[14]), rel=2e-7)


def test_irradiance_ab0_golden(G, golden):
    k = G["trace_I_ab0"]
    ctx = _lib.Context(0)
    ctx.load_octree(golden / "trace.oct")
    ctx.set_options(["-ab", "0", "-dt", "0", "-dj", "0", "-dc", "1"])
    vals, _ = ctx.rtrace(np.array(k["rays"]), flags=_lib.RB_IRRAD_RTRACE, want_results=False)
    want = np.array([[float(x) for x in ln.split()] for ln in k["out"].strip().split("\n")])
    np.testing.assert_allclose(vals, want, rtol=1e-5, atol=1e-9)


def test_config1_grid_matches_oracle(G, golden):
    """BASELINE config 1: rtrace -I -ab 0 on a 10k-point grid over the small model."""
    gx, gy = np.meshgrid(np.linspace(1, 39, 100), np.linspace(2, 45, 100))
    grid = np.stack([gx.ravel(), gy.ravel(), np.full(10000, 2.5), np.zeros(10000), np.zeros(10000), np.ones(10000)], 1)
    ctx = _lib.Context(0)
    ctx.load_octree(golden / "trace.oct")
    ctx.set_options(["-ab", "0", "-dt", "0", "-dj", "0", "-dc", "1"])
    vals, _ = ctx.rtrace(grid, flags=_lib.RB_IRRAD_RTRACE, want_results=False)
    want = port.Scene(golden / "trace.oct", dstrsrc=0.0).rtrace(grid, irrad=1)["value"]
    np.testing.assert_allclose(vals, want, rtol=1e-5, atol=1e-9)
    k = G["trace_grid_I_ab0"]
    assert int((vals[:, 0] > 0).sum()) == k["nonzero_rows"]
    assert vals.sum() == pytest.approx(k["sum"], rel=1e-6)


@pytest.mark.parametrize("name", ["bins_reinhartb_mf1", "bins_reinhartb_mf4", "bins_reinhart_mf2", "bins_klems_full",
                                  "bins_klems_half", "bins_klems_quarter", "bins_hemi", "bins_shirchiu"])
def test_bins_on_device(G, golden, name):
    """-ab 0 from above the scene: each ray lands coefficient 1 in exactly one column."""
    up = np.load(golden / "bin_dirs.npy")
    args = G[name]["args"]
    ctx = _lib.Context(0, _lib.RB_PROGRAM_RCONTRIB)
    ctx.load_octree(golden / "contrib.oct")
    ctx.set_options(["-ab", "0"])
    i, params, binv, bn = 0, "", "0", "1"
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
    assert m.shape == (3000, G[name]["ncols"], 3)
    assert np.all(m.sum(axis=1) == 1.0)
    assert np.array_equal(m[:, :, 0].argmax(axis=1), np.array(G[name]["bins"]))


@pytest.mark.parametrize("fixture,nrays", [("office2k", 100_000), ("office100k", 1_000_000)])
def test_hits_bit_exact_vs_oracle(fixture, nrays, request):
    octf = request.getfixturevalue(fixture)
    rays = scenegen.random_rays(nrays, seed=21)
    ctx = _lib.Context(0)
    ctx.load_octree(octf)
    ctx.set_options(["-ab", "0"])
    _, res = ctx.rtrace(rays, want_values=False)
    want = port.Scene(octf).rtrace(rays)
    assert np.array_equal(res["robj"], want["robj"])            # surface, bit-exact
    assert np.array_equal(res["omod"], want["omod"])            # modifier, bit-exact
    hit = want["robj"] >= 0
    np.testing.assert_allclose(res["rot"][hit], want["rot"][hit], rtol=1e-12)
    np.testing.assert_allclose(res["rop"][hit], want["rop"][hit], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(res["ron"][hit], want["ron"][hit], atol=1e-12)
    kinds = {ctx.object_type(int(i)) for i in np.unique(res["robj"][hit])}
    if fixture == "office100k":
        assert {"polygon", "sphere", "cylinder"} <= kinds        # all three intersectors exercised


@pytest.mark.skipif(not refrun.available(), reason="oracle/_ref not on this box")
def test_hits_bit_exact_vs_reference_binary(office2k):
    rays = scenegen.random_rays(50_000, seed=22)
    ctx = _lib.Context(0)
    ctx.load_octree(office2k)
    ctx.set_options(["-ab", "0"])
    _, res = ctx.rtrace(rays, want_values=False)
    s, m = names(ctx, res["robj"]), names(ctx, res["omod"])
    ref = refrun.rtrace(office2k, rays, ["-ab", "0", "-osmL"]).splitlines()
    for i, line in enumerate(ref):
        f = line.split("\t")
        assert (s[i], m[i]) == (f[0], f[1]), i
        assert res["rot"][i] == pytest.approx(float(f[2]), rel=1e-6)


def test_python_rtrace_bytes_equal_reference_text(G, golden):
    """The drop-in rtrace() prints the same ASCII as the reference did."""
    k = G["trace_ovposmNL"]
    rays = "\n".join(" ".join(str(v) for v in r) for r in k["rays"]).encode()
    out = pr.rtrace(rays, golden / "trace.oct", header=False, outspec="posmNL", params=["-ab", "0"]).decode()
    want = ["\t".join(ln.split("\t")[3:]) for ln in k["out"].strip("\n").split("\n")]
    assert out.strip("\n").split("\n") == want
    hdr = pr.rtrace(rays, golden / "trace.oct", outspec="L", params=["-ab", "0"]).decode()
    assert hdr.startswith("#?RADIANCE\n") and "FORMAT=ascii\n\n" in hdr and "NCOMP=1" in hdr
    dbl = pr.rtrace(np.array(k["rays"]).tobytes(), golden / "trace.oct", header=False, inform="d", outform="d",
                    outspec="L", params=["-ab", "0"])
    assert np.frombuffer(dbl, dtype=np.float64)[0] == pytest.approx(6.0037202, rel=1e-7)


# --------------------------------------------------------------- stochastic --
def test_unobstructed_sensor_sums_to_pi(golden):
    sens = np.array([[20, 20, 12, 0, 0, 1]], dtype=float)
    ctx = rc_ctx(golden / "contrib.oct", ["-ab", "1", "-ad", "4096", "-lw", "1e-4"])
    m = ctx.rcontrib(sens, flags=_lib.RB_IRRAD_RCONTRIB)
    assert m[0, :, 0].sum() == pytest.approx(np.pi, rel=1e-5)
    assert np.array_equal(m[..., 0], m[..., 1]) and np.array_equal(m[..., 0], m[..., 2])


def _stat_compare(a, b, reps_a, reps_b, what):
    """|mean_a - mean_b| <= 5 sigma of the difference (sigma estimated from the
    repeated runs, 16 each, so the estimate itself is good to ~20 %) + 0.3 % of
    the value, for every row sum.  The reference seeds from time(0), so this
    test sees different reference samples every run; 5 sigma keeps the
    false-alarm rate below 1e-4."""
    ma, mb = a.mean(0), b.mean(0)
    sig = np.sqrt(a.var(0, ddof=1) / reps_a + b.var(0, ddof=1) / reps_b)
    bad = np.abs(ma - mb) > 5 * sig + 3e-3 * np.abs(mb) + 1e-6
    assert not bad.any(), (what, ma[bad], mb[bad], sig[bad])


def test_stochastic_row_sums_vs_oracle_and_reference(golden):
    sens = np.array([[10, 10, 3, 0, 0, 1], [4, 5, 3, 0, 0, 1], [30, 40, 5, 0, 0, 1]], dtype=float)
    opts = ["-ab", "3", "-ad", "2048", "-lw", "1e-4"]
    reps = 16
    gpu, orc = [], []
    for k in range(reps):
        ctx = rc_ctx(golden / "contrib.oct", opts)
        ctx.set_seed(1000 + k)
        gpu.append(ctx.rcontrib(sens, flags=_lib.RB_IRRAD_RCONTRIB)[:, :, 0].sum(1))
        s = port.Scene(golden / "contrib.oct", rcontrib=True, ambounce=3, ambdiv=2048, minweight=1e-4, seed=7 + k)
        s.add_modifier("skyglow", port.BIN_REINHARTB, 1, (0, 0, -1), (0, 1, 0), 1.0, 145)
        orc.append(s.rcontrib(sens, irrad=2)[:, :, 0].sum(1))
    gpu, orc = np.array(gpu), np.array(orc)
    _stat_compare(gpu, orc, reps, reps, "gpu vs oracle")
    if refrun.available():
        # the reference seeds from time(0): repeat the sensors inside ONE run so
        # the repetitions are independent draws of its random sequence
        ref = refrun.rcontrib(golden / "contrib.oct", np.tile(sens, (reps, 1)), ["-I"] + opts + RB_ARGS)
        ref = ref.reshape(reps, 3, -1, 3)[:, :, :, 0].sum(2)
        _stat_compare(gpu, ref, reps, reps, "gpu vs reference")


def test_stochastic_bins_office_vs_oracle(office2k):
    """Per-bin agreement on a cluttered scene with glass.  Sensors stand at the
    south windows and look out, so each sees the sky through `glass` (two rays
    per pane hit) and the room by reflection.  Tolerance: the matrix total and
    every bin with >= 40 expected first-level hits must agree within 4 sigma,
    sigma^2 = the Poisson variance of the hit count of both Monte-Carlo runs
    (stratified sampling has less; two oracle runs with different seeds stay
    below 0.15 of this tolerance), plus 1 %."""
    n = 16
    x = np.linspace(2.5, 37.5, n)
    d = np.array([0.0, -1.0, 0.25]) / np.linalg.norm([0.0, -1.0, 0.25])
    sens = np.stack([x, np.full(n, 1.5), np.full(n, 1.8)] + [np.full(n, v) for v in d], axis=1)
    opts = ["-ab", "2", "-ad", "1024", "-lw", "1e-3"]
    ctx = rc_ctx(office2k, opts)
    g = ctx.rcontrib(sens, flags=_lib.RB_IRRAD_RCONTRIB).astype(np.float64)[:, :, 0]
    s = port.Scene(office2k, rcontrib=True, ambounce=2, ambdiv=1024, minweight=1e-3, seed=3)
    s.add_modifier("skyglow", port.BIN_REINHARTB, 1, (0, 0, -1), (0, 1, 0), 1.0, 145)
    o = s.rcontrib(sens, irrad=2)[:, :, 0]
    w = np.pi / 1024                      # weight of one first-level sample
    assert o.sum() > 400 * w              # the test has signal
    sig_tot = np.sqrt((g.sum() + o.sum()) * w)
    assert abs(g.sum() - o.sum()) <= 4 * sig_tot + 0.01 * o.sum()
    pg, po = g.sum(0), o.sum(0)
    nz = po > 40 * w
    assert nz.sum() >= 5
    sig = np.sqrt((pg + po) * w)
    assert np.all(np.abs(pg[nz] - po[nz]) <= 4 * sig[nz] + 0.01 * po[nz])


def test_stochastic_per_bin_vs_high_sample_reference(office2k):
    """SURVEY 8(d) parity protocol, literally: the reference run with 32x the
    samples (`-c 32` accumulation of the same sensors) is the truth c*; every
    bin of the GPU matrix with >= 30 expected first-level hits must satisfy
    |c_gpu - c*| <= 4 sqrt(sigma_gpu^2 + sigma_ref^2) for >= 99.9 % of such
    bins (sigma^2 from the Poisson variance of the hit count: stratified
    sampling has less), row sums within 1 %; and the same statistic between
    two halves of the reference's own repetitions sets the scale: the GPU's
    normalised chi^2 may not exceed twice theirs."""
    if not refrun.available():
        pytest.skip("oracle/_ref (the unmodified reference binaries) is not on this box")
    n, acc = 8, 32
    x = np.linspace(4, 36, n)
    d = np.array([0.0, -1.0, 0.3]) / np.linalg.norm([0.0, -1.0, 0.3])
    sens = np.stack([x, np.full(n, 1.5), np.full(n, 1.8)] + [np.full(n, v) for v in d], axis=1)
    opts = ["-ab", "2", "-ad", "2048", "-lw", "5e-4"]
    w = np.pi / 2048                                   # weight of one first-level sample
    rep = np.repeat(sens, acc, axis=0)                 # acc consecutive copies of each sensor
    ref_all = refrun.rcontrib(office2k, rep, ["-I"] + opts + RB_ARGS, nproc=8).reshape(n, acc, 145, 3)[..., 0]
    cstar = ref_all.mean(1)                            # 32x samples
    ctx = rc_ctx(office2k, opts)
    g = ctx.rcontrib(rep, flags=_lib.RB_IRRAD_RCONTRIB, accum=acc).astype(np.float64)[:, :, 0]     # same job on the GPU (also 32x)
    g1 = ctx.rcontrib(sens, flags=_lib.RB_IRRAD_RCONTRIB).astype(np.float64)[:, :, 0]              # and a single-sample run
    big = cstar * acc >= 30 * w                        # >= 30 expected hits in a 32x run
    assert big.sum() >= 40
    for name, c, k in (("32x", g, acc), ("1x", g1, 1)):
        sel = cstar * k >= 30 * w                      # >= 30 expected hits in THIS run
        sig = np.sqrt(cstar * w / k + cstar * w / acc)
        z = np.abs(c - cstar)[sel] / sig[sel]
        # 99.9 % of a few hundred bins = all but (at most) one; nothing beyond 6 sigma
        assert sel.sum() == 0 or ((z > 4).sum() <= max(1, int(1e-3 * sel.sum())) and z.max() <= 6), (name, z.max())
        assert np.all(np.abs(c.sum(1) - cstar.sum(1)) <= 0.01 * cstar.sum(1) + 4 * np.sqrt(cstar.sum(1) * w * (1 / k + 1 / acc))), name
    # reference vs reference: two halves of its own repetitions give the chi^2 scale
    ha, hb = ref_all[:, :acc // 2].mean(1), ref_all[:, acc // 2:].mean(1)
    chi_ref = (((ha - hb) ** 2)[big] / (2 * cstar[big] * w / (acc // 2))).mean()
    gh = ctx.rcontrib(np.repeat(sens, acc // 2, axis=0), flags=_lib.RB_IRRAD_RCONTRIB, accum=acc // 2, row_base=10_000).astype(np.float64)[:, :, 0]
    chi_gpu = (((gh - ha) ** 2)[big] / (2 * cstar[big] * w / (acc // 2))).mean()
    assert chi_gpu <= 2.0 * chi_ref + 0.2, (chi_gpu, chi_ref)


# ------------------------------------------------------------- invariances --
def test_result_independent_of_batching_and_sharding(office2k):
    sens = scenegen.office_sensors(64, seed=9)
    opts = ["-ab", "2", "-ad", "256", "-lw", "4e-3"]
    ctx = rc_ctx(office2k, opts)
    whole = ctx.rcontrib(sens, flags=_lib.RB_IRRAD_RCONTRIB, dtype=np.float64)
    again = ctx.rcontrib(sens, flags=_lib.RB_IRRAD_RCONTRIB, dtype=np.float64)
    np.testing.assert_allclose(again, whole, rtol=1e-12, atol=1e-15)      # same seed: same paths (sum order may differ)
    small = rc_ctx(office2k, opts)
    small.set_queue_capacity(8192)       # forces many batches + overflow retries
    parts = small.rcontrib(sens, flags=_lib.RB_IRRAD_RCONTRIB, dtype=np.float64)
    np.testing.assert_allclose(parts, whole, rtol=1e-12, atol=1e-15)
    assert small.stats()["batches"] > 1
    # two "ranks": rows [0,32) and [32,64) with row_base = global row index
    a = ctx.rcontrib(sens[:32], flags=_lib.RB_IRRAD_RCONTRIB, row_base=0, dtype=np.float64)
    b = ctx.rcontrib(sens[32:], flags=_lib.RB_IRRAD_RCONTRIB, row_base=32, dtype=np.float64)
    np.testing.assert_allclose(np.concatenate([a, b]), whole, rtol=1e-12, atol=1e-15)


def test_accumulate_and_edge_cases(golden):
    ctx = rc_ctx(golden / "contrib.oct", ["-ab", "0"])
    up = np.load(golden / "bin_dirs.npy")[:10]
    assert ctx.rcontrib(np.zeros((0, 6))).shape == (0, 145, 3)                  # empty input
    m1 = ctx.rcontrib(up)
    m3 = ctx.rcontrib(up, accum=3)                                              # ragged: 10 rays -> 4 records
    assert m3.shape == (4, 145, 3)
    np.testing.assert_allclose(m3[0], m1[0:3].sum(0) / 3, rtol=1e-6)
    np.testing.assert_allclose(m3[3], m1[9:10].sum(0), rtol=1e-6)      # partial record: averaged over its 1 ray
    z = up.copy()
    z[4, 3:] = 0                                                               # zero direction = blank record
    mz = ctx.rcontrib(z)
    assert mz[4].sum() == 0 and np.array_equal(mz[5], m1[5])


# -------------------------------------------------------------- rejections --
def test_explicit_rejections(workdir, golden):
    rad = workdir / "bad.rad"
    rad.write_text(scenegen.MATERIALS + scenegen.SKY +
                   "void mirror crystal\n0\n0\n3 .9 .9 .9\n\n"
                   "crystal polygon slab\n0\n0\n12 0 0 1  4 0 1  4 4 1  0 4 1\n\n"
                   "void light lamp\n0\n0\n3 10 10 10\n\n")
    octf = workdir / "bad.oct"
    scenegen.build_octree(rad, octf)
    ctx = _lib.Context(0)
    ctx.load_octree(octf)
    ctx.set_options(["-ab", "0"])
    with pytest.raises(_lib.RBError, match="unsupported material.*mirror"):
        ctx.rtrace(np.array([[2, 2, 0, 0, 0, 1.0]]))
    ctx.rtrace(np.array([[9, 9, 0, 0, 0, 1.0]]))              # a ray that never meets it is fine
    rad6 = workdir / "brushed.rad"                            # plastic2 orientation given as a .cal expression
    rad6.write_text(scenegen.MATERIALS + scenegen.SKY + "void plastic2 brushed\n4 Ny -Nx 0 .\n0\n6 .5 .5 .5 .1 .1 .3\n\n"
                    "brushed polygon plate\n0\n0\n12 0 0 1  4 0 1  4 4 1  0 4 1\n\n")
    oct6 = workdir / "brushed.oct"
    scenegen.build_octree(rad6, oct6)
    ctx6 = _lib.Context(0)
    ctx6.load_octree(oct6)
    ctx6.set_options(["-ab", "0"])
    with pytest.raises(_lib.RBError, match="unsupported material.*plastic2"):
        ctx6.rtrace(np.array([[2, 2, 3, 0, 0, -1.0]]))
    ctx2 = _lib.Context(0)
    ctx2.load_octree(golden / "trace.oct")
    ctx2.set_options(["-ab", "2"])                            # rtrace default -aa .1
    with pytest.raises(_lib.RBError, match="irradiance cache"):
        ctx2.rtrace(np.array([[2, 2, 0, 0, 0, 1.0]]))
    ctx2.set_options(["-ab", "1", "-aa", "0", "-ad", "2048"])  # rtrace default -as 512: ambsupersamp() would run
    with pytest.raises(_lib.RBError, match="super-sampling.*-as 0"):
        ctx2.rtrace(np.array([[2, 2, 0, 0, 0, 1.0]]))
    ctx2.set_options(["-ad", "16"])                           # 16 divisions never super-sample (MINADIV, ambcomp.c:413)
    ctx2.rtrace(np.array([[2, 2, 0, 0, 0, 1.0]]))
    ctx2.set_options(["-ab", "0", "-lr", "80"])
    with pytest.raises(_lib.RBError, match="-lr beyond"):
        ctx2.rtrace(np.array([[2, 2, 0, 0, 0, 1.0]]))
    ctx7 = _lib.Context(0)                                    # rtrace's default -dt .03 on a scene with many lamps:
    ctx7.load_octree(golden / "lights" / "lights.oct")        # adaptive shadow testing is not built -> by name
    ctx7.set_options(["-ab", "0"])
    with pytest.raises(_lib.RBError, match="-dt 0.03.*pass -dt 0"):
        ctx7.rtrace(np.array([[2, 2, 1, 0, 0, -1.0]]))
    ctx7.set_options(["-dt", "0"])
    ctx7.rtrace(np.array([[2, 2, 1, 0, 0, -1.0]]))
    ctx8 = _lib.Context(0)                                    # one sun + a glow sky: at most MINSHADCNT candidates,
    ctx8.load_octree(golden / "trace.oct")                    # the threshold never acts (source.c:490) -> accepted
    ctx8.set_options(["-ab", "0", "-dt", ".5"])
    a, _ = ctx8.rtrace(np.array([[20, 20, 9.5, 0, 0, 1.0]]), flags=_lib.RB_IRRAD_RTRACE)
    ctx8.set_options(["-dt", "0"])
    b, _ = ctx8.rtrace(np.array([[20, 20, 9.5, 0, 0, 1.0]]), flags=_lib.RB_IRRAD_RTRACE)
    assert np.array_equal(a, b) and a[0, 0] > 100
    rad5 = workdir / "dirtsky.rad"                            # only the two sky .cal functions are native code
    rad5.write_text("void brightfunc dirt\n2 dirtfn dirt.cal\n0\n0\n\ndirt glow skyglow\n0\n0\n4 1 1 1 0\n\n"
                    "skyglow source sky\n0\n0\n4 0 0 1 180\n\n")
    oct5 = workdir / "dirtsky.oct"
    scenegen.build_octree(rad5, oct5)
    with pytest.raises(_lib.RBError, match="brightfunc|unsupported modifier"):
        ctx3 = _lib.Context(0)
        ctx3.load_octree(oct5)
        ctx3.set_options(["-ab", "0"])
        ctx3.rtrace(np.array([[4, 5, 6, 0, 0, 1.0]]))         # value under an arbitrary .cal pattern is not built
    rad2 = workdir / "lamp.rad"
    rad2.write_text(scenegen.MATERIALS + "void light lamp\n0\n0\n3 10 10 10\n\n"
                    "lamp polygon fixture\n0\n0\n12 0 0 3  0 1 3  1 1 3  1 0 3\n\n"
                    "floor_mat polygon fl\n0\n0\n12 0 0 0  4 0 0  4 4 0  0 4 0\n\n")
    oct2 = workdir / "lamp.oct"
    scenegen.build_octree(rad2, oct2)
    ctx4 = _lib.Context(0)
    ctx4.load_octree(oct2)                                    # local emitters are built (test_local_light_sources)
    ctx4.set_options(["-ab", "0", "-dt", "0"])                # (a lamp may split into partitions: -dt > 0 is rejected)
    v, _ = ctx4.rtrace(np.array([[2, 2, 1, 0, 0, -1.0]]))
    assert v[0, 0] > 0
    rad3 = workdir / "conelamp.rad"
    rad3.write_text("void light lamp\n0\n0\n3 10 10 10\n\nlamp cone shade\n0\n0\n8 0 0 3  0 0 2.5  .1 .4\n\n")
    oct3 = workdir / "conelamp.oct"
    scenegen.build_octree(rad3, oct3)
    with pytest.raises(_lib.RBError, match="cannot be a light source"):
        _lib.Context(0).load_octree(oct3)                     # the reference: "illegal material"


# ------------------------------------------------------ Python boundaries ---
def test_rcontrib_forces_its_overrides_and_contributions_of_black_emitters(golden, workdir):
    """rcontrib forces -dt 0 -as 0 -aa 0 whatever the caller set (rcmain.c:164-171, rxcmain.cpp:154-156): a
    context carrying rtrace's defaults (-aa .1 -as 512 -dt .03) must run, with the result of the forced
    values.  -V+ on a tracked emitter seen from behind multiplies by rcol = 0 and adds nothing
    (rcontrib.c:296-301); on a tracked NON-emitter it fails by name (no returned value in a forward engine)."""
    up = np.load(golden / "bin_dirs.npy")[:200]
    want = rc_ctx(golden / "contrib.oct", ["-ab", "0"]).rcontrib(up)
    c = _lib.Context(0, _lib.RB_PROGRAM_RTRACE)               # rtrace defaults: -aa .1 -as 512 -dt .03
    c.load_octree(golden / "contrib.oct")
    c.set_options(["-ab", "1", "-ad", "2048", "-lw", "1e-4"])
    c.cal_load("reinhartb.cal"); c.cal_set(RB_P)
    c.add_modifier("skyglow", RB_P, "rbin", 145)
    m = c.rcontrib(np.array([[20, 20, 12, 0, 0, 1.0]]), flags=_lib.RB_IRRAD_RCONTRIB)
    assert abs(float(m[0, :, 0].sum()) - np.pi) < 1e-4
    c.set_options(["-ab", "0"])
    assert np.array_equal(c.rcontrib(up), want)
    rad = workdir / "backglow.rad"
    rad.write_text(scenegen.MATERIALS + scenegen.SKY + "void glow winglow\n0\n0\n4 2 3 4 0\n\n"
                   "winglow polygon win\n0\n0\n12 0 0 1  4 0 1  4 4 1  0 4 1\n\n"         # normal +z
                   "floor_mat polygon slab\n0\n0\n12 6 0 1  10 0 1  10 4 1  6 4 1\n\n")
    octf = workdir / "backglow.oct"
    scenegen.build_octree(rad, octf)
    v = _lib.Context(0, _lib.RB_PROGRAM_RCONTRIB)
    v.load_octree(octf)
    v.set_options(["-ab", "0"])
    v.add_modifier("winglow", "", "0", 1)
    rays = np.array([[2, 2, 3, 0, 0, -1.0], [2, 2, 0, 0, 0, 1.0]])       # front, then from behind
    cm = v.rcontrib(rays, flags=_lib.RB_FLAG_CONTRIB, dtype=np.float64)
    np.testing.assert_allclose(cm[0, 0], [2, 3, 4], rtol=1e-6)
    assert not cm[1].any()
    v.add_modifier("floor_mat", "", "0", 1)
    with pytest.raises(_lib.RBError, match="-V\\+.*does not emit"):
        v.rcontrib(np.array([[8, 2, 3, 0, 0, -1.0]]), flags=_lib.RB_FLAG_CONTRIB)
    assert v.rcontrib(np.array([[8, 2, 3, 0, 0, -1.0]]))[0, 1, 0] == 1.0     # coefficients (-V-) are fine


def test_rcontrib_class_matches_reference_layout(golden):
    rays = b"10 10 3 0 0 1\n4 5 3 0 0 1\n"
    rc = pr.Rcontrib(rays, golden / "contrib.oct", yres=2, params=["-I", "-ab", "0"])
    rc.add_modifier("skyglow", binv="if(-Dx*0-Dy*0-Dz*-1,0,-1)", nbins="1").add_modifier("groundglow", binv="0")
    out = rc().decode()
    head, body = out.split("\n\n", 1)
    assert "NROWS=2" in head and "NCOLS=2" in head and "NCOMP=3" in head and "FORMAT=ascii" in head
    rows = body.strip("\n").split("\n")
    assert len(rows) == 2 and all(len(r.split("\t")) == 7 for r in rows)      # 2 cols x 3 + trailing tab
    rc2 = pr.Rcontrib(np.array([[20, 20, 12, 0, 0, 1.0]]).tobytes(), golden / "contrib.oct", inform="d", outform="f",
                      params=["-I", "-ab", "1", "-ad", "1024", "-lw", "1e-3", "-h"])
    rc2.add_modifier("skyglow", calfile="reinhartb.cal", param=RB_P, nbins="Nrbins", binv="rbin")
    m = np.frombuffer(rc2(), dtype=np.float32).reshape(1, 145, 3)
    assert m[0, :, 0].sum() == pytest.approx(np.pi, rel=1e-5)
    with pytest.raises(RuntimeError, match="unsupported function file"):
        pr.Rcontrib(rays, golden / "contrib.oct").add_modifier("skyglow", calfile="tregenza.cal", binv="tbin")()


def test_rcontrib_simul_manager_like_reference_test(golden, workdir):
    """Mirror of /root/reference/tests/test_rcontrib.py:9-75, then what the reference's managers add around it:
    the header calls (radiance_ext.cpp:305-317), the output file the two modifiers share (a Radiance matrix file,
    RcontribSimulManager.cpp:452-519) and RcOutputOp.RECOVER (:426-438, 537-596)."""
    outfile = str(workdir / "test.mtx")
    rays = np.array([[10, 10, 3], [0.0, 0.0, 1.0], [4.0, 5.0, 3.0], [0.0, 0.0, 1.0]])
    pr.initfunc()
    pr.calcontext(pr.RCCONTEXT)
    rp = pr.get_ray_params()
    rp.u = True; rp.dj = 0.9; rp.dr = 3; rp.dp = 512; rp.ds = 0.2; rp.st = 0.02; rp.ss = 1; rp.lr = -10
    rp.lw = 2e-3; rp.ar = 256; rp.ad = 350; rp.ab = 6; rp.aa = 0; rp.as_ = 0; rp.st = 0
    mgr = pr.RcontribSimulManager()
    mgr.accum = 1
    pr.loadfunc("reinhartb.cal")
    pr.set_eparams(RB_P)
    bincnt = int(pr.eval("Nrbins") + 0.5)
    assert bincnt == 145
    mgr.yres = rays.shape[0] // 2
    mgr.set_flag(pr.RTimmIrrad, True)
    mgr.add_modifier(modn="groundglow", outspec=outfile, binval="if(-Dx*0-Dy*0-Dz*-1,0,-1)", bincnt=1)
    mgr.add_modifier(modn="skyglow", outspec=outfile, prms=RB_P, binval="rbin", bincnt=bincnt)
    out = mgr.get_output()
    assert out.get_name() == outfile and out.row_bytes == 146 * 3 * 4
    pr.set_ray_params(rp)
    mgr.load_octree(str(golden / "contrib.oct"))
    assert "oconv" in mgr.get_head_str() and mgr.get_head_len() == len(mgr.get_head_str())
    mgr.add_header(["rcontrib", "-ab", "6", "a b"])
    mgr.add_header("SOFTWARE= test")
    assert mgr.get_head_str("SOFTWARE=") == " test" and mgr.get_head_str("NOPE=") is None
    assert mgr.get_format() == ord("f")
    mgr.out_op = pr.RcOutputOp.FORCE
    assert mgr.prep_output() == 0                # rows already there
    mgr.set_thread_count(1)
    mgr.rcontrib(rays)
    result = np.array(mgr.get_output_array())
    assert result.shape == (2, 146 * 3) and result.dtype == np.float32
    assert result.sum() > 5.0                    # the reference's own assertion
    mgr.cleanup(False)
    raw = open(outfile, "rb").read()
    head, _, body = raw.partition(b"\n\n")
    assert head.startswith(b"#?RADIANCE\n") and b'rcontrib -ab 6 "a b"\n' in head and b"NROWS=0000000000000002\n" in head
    assert b"NCOLS=146\nNCOMP=3\nBigEndian=0\nFORMAT=float" in head and (len(head) + 2) % 4 == 0
    assert np.array_equal(np.frombuffer(body, dtype=np.float32).reshape(2, -1), result)
    if refrun.available():                       # the file is a matrix any Radiance tool reads
        txt = refrun.run("rmtxop", ["-fa", outfile]).decode()
        assert "NROWS=2" in txt and "NCOLS=146" in txt
    with pytest.raises(RuntimeError, match="file exists"):
        mgr.out_op = pr.RcOutputOp.NEW
        mgr.prep_output()
    # RECOVER: pretend the job died after the first row; the second row is computed into the same file
    with open(outfile, "r+b") as f:
        f.seek(raw.index(b"NROWS=") + 6)
        f.write(b"%016d" % 1)
        f.seek(len(head) + 2 + 146 * 12)
        f.write(b"\0" * (146 * 12))
    mgr.out_op = pr.RcOutputOp.RECOVER
    assert mgr.prep_output() == 1 and mgr.get_row_count() == 1
    mgr.compute_record(rays[2:4])
    assert mgr.get_row_finished() == 2
    rec = np.array(mgr.get_output_array())
    assert np.array_equal(rec[0], result[0])
    assert rec[1].sum() == pytest.approx(result[1].sum(), rel=0.25) and rec[1].sum() > 0      # a new stochastic estimate of row 1
    mgr.cleanup(True)
    assert b"NROWS=0000000000000002\n" in open(outfile, "rb").read(600)
    with pytest.raises(RuntimeError, match="one file per bin"):
        pr.RcontribSimulManager().add_modifier("skyglow", "bin%d.mtx", binval="rbin", bincnt=145)
    pr.set_ray_params(None)


def test_rtrace_simul_manager_like_reference_test(golden):
    """Mirror of /root/reference/tests/test_rtrace.py:48-74."""
    got = []
    rays = np.array([[1.0, 2.0, 3.0], [0.0, 0.0, 1.0], [4.0, 5.0, 6.0], [0.0, 1.0, 0.0]])
    rp = pr.get_ray_params()
    rp.ab = 0
    pr.set_ray_params(rp)
    mgr = pr.RtraceSimulManager()
    mgr.load_octree(str(golden / "trace.oct"))
    mgr.set_thread_count(1)
    mgr.set_cooked_call(lambda ray, cd: got.append(("cooked", ray.rop)) or 0)
    mgr.set_trace_call(lambda ray, cd: got.append(("trace", ray.rop)) or 0)
    mgr.rt_flags = pr.RTdoFIFO
    assert mgr.enqueue_bundle(rays) == 2         # ray 2 sees the perezlum.cal sky: its value is native code now
    mgr.flush_queue()
    assert len([g for g in got if g[0] == "cooked"]) == 2
    mgr.cleanup_callbacks()
    assert mgr.enqueue_bundle(rays) == 2
    mgr.flush_queue()
    mgr.cleanup(True)


# ---------------------------------------------- more of the rcontrib boundary --
def test_output_files_and_accumulate_zero(golden, workdir):
    """-o spec with %s / %d (rc2.c:40-92,150-254), -fo, and -c 0 (sum of all rays)."""
    import os
    up = np.load(golden / "bin_dirs.npy")[:50]
    rays = up.tobytes()
    cwd = os.getcwd()
    os.chdir(workdir)
    try:
        rc = pr.Rcontrib(rays, golden / "contrib.oct", inform="d", outform="d", params=["-ab", "0"])
        rc.add_modifier("skyglow", calfile="klems_quarter.cal", nbins="Nkqbins", binv="kqbin(0,0,-1,0,1,0)",
                        output="q_%s.dat")
        rc.add_modifier("groundglow", binv="0", output="g_%s_%d.dat")
        assert rc() == b""                                   # everything went to files
        q = (workdir / "q_skyglow.dat").read_bytes()
        head, body = q.split(b"\n\n", 1)
        assert b"MODIFIER=skyglow" in head and b"NCOLS=41" in head and b"FORMAT=double" in head
        m = np.frombuffer(body, dtype=np.float64).reshape(50, 41, 3)
        assert np.all(m.sum(axis=1) == 1.0)
        gfile = (workdir / "g_groundglow_0.dat").read_bytes()
        assert b"MODIFIER=groundglow" in gfile and b"BIN=0" in gfile
        with pytest.raises(RuntimeError, match="cannot open 'q_skyglow.dat' for writing"):
            rc()                                             # refuses to overwrite ...
        rc.cmd.insert(1, "-fo")
        rc()                                                 # ... unless -fo
    finally:
        os.chdir(cwd)
    rc0 = pr.Rcontrib(rays, golden / "contrib.oct", inform="d", outform="d", params=["-ab", "0", "-c", "0", "-h"])
    rc0.add_modifier("skyglow", calfile="klems_quarter.cal", nbins="Nkqbins", binv="kqbin(0,0,-1,0,1,0)")
    tot = np.frombuffer(rc0(), dtype=np.float64).reshape(1, 41, 3)
    np.testing.assert_allclose(tot[0], m.sum(axis=0), rtol=1e-12)
    if refrun.available():
        ref = refrun.rcontrib(golden / "contrib.oct", up, ["-ab", "0", "-c", "0", "-f", "klems_quarter.cal", "-bn", "Nkqbins",
                                                           "-b", "kqbin(0,0,-1,0,1,0)", "-m", "skyglow"]).reshape(1, 41, 3)
        np.testing.assert_allclose(tot, ref, rtol=1e-12)


def test_view_matrix_klems_window_groups(workdir):
    """BASELINE config 4 in miniature: Klems full bins per window group, 8 glow
    window modifiers tracked at once, view rays from inside the room; per-group
    totals against the reference / oracle within Monte-Carlo tolerance."""
    import io
    rng = np.random.default_rng(5)
    out = io.StringIO()
    out.write(scenegen.MATERIALS)
    for i in range(8):
        out.write(f"void glow wg{i}\n0\n0\n4 1 1 1 0\n\n")
    buf = io.StringIO()
    scenegen.office_floor(buf, rng, 0.0, 300, tag="f0")
    txt = buf.getvalue()
    for i in range(8):
        txt = txt.replace(f"win_glass polygon f0.win{i}\n", f"wg{i} polygon f0.win{i}\n")
    out.write(txt)
    rad, octf = workdir / "view.rad", workdir / "view.oct"
    rad.write_text(out.getvalue())
    scenegen.build_octree(rad, octf)
    rays = scenegen.random_rays(4000, seed=9, lo=(5, 3, 0.9), hi=(35, 12, 2.5))
    rays[:, 4] = -np.abs(rays[:, 4]) - 0.3                   # look towards the south facade
    rays[:, 3:6] /= np.linalg.norm(rays[:, 3:6], axis=1, keepdims=True)
    opts = ["-ab", "2", "-ad", "512", "-lw", "2e-3"]
    ctx = _lib.Context(0, _lib.RB_PROGRAM_RCONTRIB)
    ctx.load_octree(octf)
    ctx.set_options(opts)
    ctx.cal_load("klems_full.cal")
    s = port.Scene(octf, rcontrib=True, ambounce=2, ambdiv=512, minweight=2e-3, seed=11)
    args = ["-f", "klems_full.cal", "-p", "RHS=+1", "-bn", "Nkbins", "-b", "kbinS"]
    for i in range(8):
        ctx.add_modifier(f"wg{i}", "RHS=+1", "kbinS", int(ctx.cal_eval("Nkbins") + .5))
        s.add_modifier(f"wg{i}", port.BIN_KLEMS_FULL, 1, (0, 1, 0), (0, 0, 1), 1.0, 145)
        args += ["-m", f"wg{i}"]
    assert ctx.num_columns() == 8 * 145
    g = ctx.rcontrib(rays).astype(np.float64)[:, :, 0].reshape(-1, 8, 145)
    o = s.rcontrib(rays)[:, :, 0].reshape(-1, 8, 145)
    # deterministic part: primary rays that see a window directly land coefficient 1 in one bin
    direct_g = (g == 1.0).sum()
    assert direct_g > 100 and direct_g == (o == 1.0).sum()
    assert np.array_equal(np.argwhere(g == 1.0), np.argwhere(o == 1.0))
    tg, to = g.sum((0, 2)), o.sum((0, 2))
    assert np.all(np.abs(tg - to) <= 0.03 * to + 0.5)       # per window group, Monte-Carlo part included
    if refrun.available():
        r = refrun.rcontrib(octf, rays, opts + args, nproc=4).reshape(-1, 8, 145, 3)[..., 0]
        assert np.array_equal(np.argwhere(r == 1.0), np.argwhere(g == 1.0))
        assert np.all(np.abs(tg - r.sum((0, 2))) <= 0.03 * to + 0.5)


def test_full_size_properties(office100k):
    """BASELINE config 2 at full scene size on a slice of the sensors:
    size-independent properties of a daylight-coefficient matrix."""
    sens = scenegen.office_sensors(100_000)[::50]            # 2000 of the 100k sensors
    ctx = rc_ctx(office100k, ["-ab", "3", "-ad", "4096", "-lw", f"{1 / 4096:.4e}"])
    m = ctx.rcontrib(sens, flags=_lib.RB_IRRAD_RCONTRIB, dtype=np.float64)
    st = ctx.stats()
    assert m.shape == (2000, 145, 3) and np.isfinite(m).all() and (m >= 0).all()
    assert m[..., 0].sum(1).max() <= np.pi * (1 + 1e-9)      # a sensor cannot see more than the whole sky
    # every material of the scene has R >= G >= B, so every coefficient product has too
    assert m.sum() > 0 and (m[..., 0] >= m[..., 1] - 1e-12).all() and (m[..., 1] >= m[..., 2] - 1e-12).all()
    assert 1.0e4 < st["nrays"] / 2000 < 1.6e4                 # ~12.6k rays per sensor (12 100)
    # records are independent: a permuted / re-based run of a subset reproduces the same rows
    sub = ctx.rcontrib(sens[500:600], flags=_lib.RB_IRRAD_RCONTRIB, row_base=500, dtype=np.float64)
    np.testing.assert_allclose(sub, m[500:600], rtol=1e-12, atol=1e-15)


@pytest.mark.skipif(not refrun.available(), reason="oracle/_ref not built")
def test_full_size_per_bin_vs_reference(office100k):
    """BASELINE configs[1] AS CONFIGURED (the 100 k-polygon office, -ab 3 -ad 4096 -lw 1/4096, Reinhart MF:1) on a
    stratified sample of its sensors, per bin against the unmodified reference on the same sensors: the statistical
    protocol of SURVEY 8(d) with the reference's own repetitions as the truth (8 repetitions of each sensor, `-c 8`
    accumulation of the same rays, on both sides).  Per bin with >= 30 expected first-level hits: within 4 sigma of
    the two estimates for all but 0.1 %, nothing beyond 6 sigma; row sums within 1 % + 4 sigma."""
    import os
    n, acc = 24, 8
    allsens = scenegen.office_sensors(100_000)
    sens = allsens[((np.arange(n) + 0.5) * len(allsens) / n).astype(int)]
    opts = ["-ab", "3", "-ad", "4096", "-lw", f"{1 / 4096:.4e}"]
    w = np.pi / 4096
    rep = np.repeat(sens, acc, axis=0)
    ref = refrun.rcontrib(office100k, rep, ["-I+", "-c", str(acc)] + opts + RB_ARGS, nproc=os.cpu_count() or 8).reshape(n, 145, 3)[..., 0]
    ctx = rc_ctx(office100k, opts)
    g = ctx.rcontrib(rep, flags=_lib.RB_IRRAD_RCONTRIB, accum=acc).astype(np.float64)[:, :, 0]
    cstar = 0.5 * (ref + g)                          # pooled estimate of the expectation (both are acc-x runs)
    sel = cstar * acc >= 30 * w                      # (deep-plan sensors see little sky: a few dozen such bins)
    assert sel.sum() >= 10, int(sel.sum())
    sig = np.sqrt(2 * cstar * w / acc)               # difference of two independent acc-x estimates
    z = np.abs(g - ref)[sel] / sig[sel]
    assert (z > 4).sum() <= max(1, int(1e-3 * sel.sum())) and z.max() <= 6, (z.max(), int((z > 4).sum()), int(sel.sum()))
    # chi^2 over every bin with >= 5 expected hits: the normalised squared differences average to ~1 (they are a sum
    # of weighted samples, not pure counts, hence the slack)
    sel5 = cstar * acc >= 5 * w
    chi = (((g - ref) ** 2)[sel5] / (sig[sel5] ** 2)).mean()
    assert sel5.sum() >= 50 and chi <= 2.0, (int(sel5.sum()), chi)
    rs_g, rs_r = g.sum(1), ref.sum(1)
    assert np.all(np.abs(rs_g - rs_r) <= 0.01 * rs_r + 4 * np.sqrt(2 * np.maximum(rs_r, rs_g) * w / acc)), (rs_g, rs_r)
    assert abs(rs_g.sum() - rs_r.sum()) <= 0.02 * rs_r.sum()


def test_instance_expansion_limit_is_refused_by_name(golden, monkeypatch):
    """Instances and meshes are flattened at load: the memory that costs is bounded, and a scene beyond
    the bound is refused naming the limit -- nested traversal (o_instance.c:16-71) is not built."""
    monkeypatch.setenv("RB_MAX_EXPANDED_SURFACES", "3")
    ctx = _lib.Context(0)
    with pytest.raises(_lib.RBError, match="RB_MAX_EXPANDED_SURFACES.*nested instance traversal is not built"):
        ctx.load_octree(golden / "volumes" / "room.oct")
    monkeypatch.delenv("RB_MAX_EXPANDED_SURFACES")
    _lib.Context(0).load_octree(golden / "volumes" / "room.oct")


@pytest.mark.parametrize("name", ["room", "meshroom"])
def test_instances_and_meshes_vs_reference_golden(G, golden, name):
    """Config-5 ingredients: octree instances and triangle meshes, against the
    reference rtrace's answers for the same rays (tests/golden/make_golden.py)."""
    ctx = _lib.Context(0)
    ctx.load_octree(golden / "volumes" / f"{name}.oct")
    ctx.set_options(["-ab", "0"])
    rays = np.load(golden / "volume_rays.npy")
    _, res = ctx.rtrace(rays, want_values=False)
    s, m = names(ctx, res["robj"]), names(ctx, res["omod"])
    for i, line in enumerate(G["volumes_" + name].splitlines()):
        f = line.split("\t")
        assert (s[i], m[i] if res["robj"][i] >= 0 else "*") == (f[0], f[1]), (i, f)
        if f[0] != "*":
            assert res["rot"][i] == pytest.approx(float(f[2]), rel=2e-6)
            np.testing.assert_allclose(res["ron"][i], [float(x) for x in f[3:6]], atol=2e-5)
    # and a daylight-coefficient run over it works end to end
    rc = _lib.Context(0, _lib.RB_PROGRAM_RCONTRIB)
    rc.load_octree(golden / "volumes" / f"{name}.oct")
    rc.set_options(["-ab", "1", "-ad", "256", "-lw", "1e-3"])
    rc.add_modifier("wall", "", "0", 1)
    m1 = rc.rcontrib(np.array([[0, 0, 3.0, 0, 0, -1.0]]))
    assert m1.shape == (1, 1, 3) and m1[0, 0, 0] > 0


def test_view_rays_on_device_equal_host_vwrays(monkeypatch):
    """vwrays on the device (C ABI rb_view_rays, csrc/rb_views.cu) against the numpy restatement
    of util/vwrays.c + common/image.c (which tests/test_host.py pins to the reference binary): all six view types,
    offsets, fore / aft planes (directions scaled by the aft distance), pixels without a ray, several rays per
    pixel; then straight into device memory and traced from there."""
    from pyradiance_b200 import views
    cases = [["-vtv", "-vp", "1", "2", "1.5", "-vd", ".3", "1", "-.1", "-vu", "0", "0", "1", "-vh", "60", "-vv", "40"],
             ["-vtl", "-vp", "5", "5", "9", "-vd", "0", "0", "-1", "-vu", "0", "1", "0", "-vh", "12", "-vv", "8", "-vo", ".5", "-va", "6"],
             ["-vtc", "-vp", "1", "2", "1.5", "-vd", "0", "1", "0", "-vu", "0", "0", "1", "-vh", "200", "-vv", "50", "-vs", ".1", "-vl", "-.2"],
             ["-vth", "-vp", "1", "2", "1.5", "-vd", "0", "0", "1", "-vu", "0", "1", "0", "-vh", "180", "-vv", "180"],
             ["-vta", "-vp", "1", "2", "1.5", "-vd", "0", "-1", "0", "-vu", "0", "0", "1", "-vh", "180", "-vv", "180", "-vo", ".1", "-va", "30"],
             ["-vts", "-vp", "1", "2", "1.5", "-vd", "1", "0", "0", "-vu", "0", "0", "1", "-vh", "160", "-vv", "120"]]
    for view in cases:
        for rep in (1, 3):
            monkeypatch.setenv("RB_VWRAYS_HOST", "1")
            host = np.frombuffer(pr.vwrays(outform="d", ray_count=rep, xres=97, yres=64, view=view), dtype=np.float64).reshape(-1, 6)
            monkeypatch.delenv("RB_VWRAYS_HOST")
            dev = np.frombuffer(pr.vwrays(outform="d", ray_count=rep, xres=97, yres=64, view=view), dtype=np.float64).reshape(-1, 6)
            assert dev.shape == host.shape and dev.shape[0] > 1000
            np.testing.assert_allclose(dev, host, rtol=0, atol=2e-12, err_msg=" ".join(view))
    # into device memory, and traced from there: the 3-phase view matrix never holds its rays on the host
    v, xr, yr = views.view_from_args(cases[3], 64, 64)
    ctx = _lib.Context(0)
    n = xr * yr
    d_rays = ctx.device_alloc(n * 48)
    assert ctx.view_rays(v, xr, yr, out_ptr=d_rays) == n
    back = np.empty((n, 6))
    ctx.device_download(back, d_rays)
    np.testing.assert_array_equal(back, ctx.view_rays(v, xr, yr))
    ctx.device_free(d_rays)
    # jitter: inside the pixel, different per ray, reproducible per seed
    j1 = ctx.view_rays(v, xr, yr, repeat=2, pj=1.0, seed=5)
    j2 = ctx.view_rays(v, xr, yr, repeat=2, pj=1.0, seed=5)
    j3 = ctx.view_rays(v, xr, yr, repeat=2, pj=1.0, seed=6)
    assert np.array_equal(j1, j2) and not np.array_equal(j1, j3) and not np.array_equal(j1[0::2], j1[1::2])


def test_device_octree_build_equals_host_build(workdir, monkeypatch):
    """the level-by-level octree build on the GPU (rb_octbuild_gpu.cu: every surface / child-cube
    overlap test of the reference, ot/o_face.c, ot/sphere.c, ot/o_cone.c, as device code) gives the file the
    threaded host builder gives -- which tests/test_host.py pins byte for byte to the reference `oconv -f`."""
    import hashlib
    import time
    for npoly, seed in ((30_000, 5), (200_000, 6)):
        rad = workdir / f"devoct{npoly}.rad"
        scenegen.write_office(rad, npolys=npoly, seed=seed)
        monkeypatch.setenv("RB_OCTBUILD_DEVICE_STRICT", "1")
        t = time.time(); _lib.oconv_file(rad, workdir / "dev.oct"); t_dev = time.time() - t
        monkeypatch.setenv("RB_OCTBUILD_HOST", "1")
        t = time.time(); _lib.oconv_file(rad, workdir / "host.oct"); t_host = time.time() - t
        monkeypatch.delenv("RB_OCTBUILD_HOST")
        a, b = (workdir / "dev.oct").read_bytes(), (workdir / "host.oct").read_bytes()
        assert hashlib.sha256(a).hexdigest() == hashlib.sha256(b).hexdigest(), (npoly, len(a), len(b))
        print(f"octree of {npoly} polygons: device path {t_dev:.2f} s, host path {t_host:.2f} s (parse and write included)")
    monkeypatch.setenv("RB_OCTBUILD_DEVICE_STRICT", "1")          # other octree parameters: -n 3 -r 2048
    rad = workdir / "devoct30000.rad"
    _lib.oconv_file(rad, workdir / "dev2.oct", objlim=3, maxres=2048)
    monkeypatch.setenv("RB_OCTBUILD_HOST", "1")
    _lib.oconv_file(rad, workdir / "host2.oct", objlim=3, maxres=2048)
    assert (workdir / "dev2.oct").read_bytes() == (workdir / "host2.oct").read_bytes()


GEOM_CASES = [("curved", "curved.oct", "curved_rays"), ("curvedtext", "curved_text.oct", "curved_rays"),
              ("coinc", "coinc.oct", "coinc_rays"), ("coincfine", "coinc_fine.oct", "coinc_rays")]


@pytest.mark.parametrize("tag,octf,rk", GEOM_CASES)
def test_cone_family_and_rayreject_vs_reference_golden(golden, monkeypatch, tag, octf, rk):
    """SURVEY 8a a5 / a7 / a8 on the device: cone, cup, cylinder, tube, ring, sphere, bubble (inside and
    outside hits, rays across the end-cap rims: `cand_other`, the cone normal of `hit_frame`) and
    rayreject()'s tie rules on coincident surfaces (also with the surfaces spread over many leaves,
    where this engine re-tests them), against the unmodified reference rtrace
    (tests/golden/make_golden_geom.py).  Surface and modifier names exact, distance 1e-9, normal 1e-9,
    -ab 0 value 1e-5.  `curvedtext` loads the NOT frozen octree (readoct.c:90-100: scene read as text)."""
    g = np.load(golden / "geom.npz")
    monkeypatch.chdir(golden / "geom")
    ctx = _lib.Context(0)
    ctx.load_octree(octf)
    ctx.set_options([str(a) for a in g["args"]])
    v, res = ctx.rtrace(g[rk])
    surf, mod = np.array(names(ctx, res["robj"])), np.array(names(ctx, res["omod"]))
    same = surf == g[tag + "_surf"]
    # The rim rays are aimed AT the edge of a cap / ring (offsets down to 0 and +-1e-9): there `a > r1*r1` is decided
    # by the last bit of a sum whose order the reference's -ffast-math build is free to change.  With the 31-bit
    # reals of the frozen octrees every such ray falls on the reference's side; with the full-precision reals of
    # the text octree one of the 6252 (ray 6229: outer edge of the tilted ring, offset 0) does not.
    assert (~same).sum() <= (2 if tag == "curvedtext" else 0), [(i, surf[i], g[tag + "_surf"][i]) for i in np.flatnonzero(~same)[:10]]
    assert np.array_equal(mod[same], g[tag + "_mod"][same])
    loc = (g[tag + "_dist"] < 1e9) & same
    v, g_value = v[same], g[tag + "_value"][same]
    np.testing.assert_allclose(res["rot"][loc], g[tag + "_dist"][loc], rtol=1e-9)
    np.testing.assert_allclose(res["ron"][loc], g[tag + "_norm"][loc], atol=1e-9)     # -oN: flips undone (rtrace.c:787-804)
    bad = ~np.isclose(v, g_value, rtol=1e-5, atol=1e-9).all(1)
    assert bad.sum() <= (3 if tag == "curvedtext" else 0), (int(bad.sum()), np.flatnonzero(same)[bad][:10], surf[same][bad][:10],
                                                            v[bad][:5], g_value[bad][:5])


def test_shadow_rays_any_blocker_is_result_neutral(golden, monkeypatch):
    """k_trace ends a shadow ray towards a DISTANT source at the first surface that is opaque to shadow rays instead of
    its nearest hit: the coefficient matrix must not depend on that -- same random keys, so the matrices of
    the miniature sun scene (145 suns, louvre instances, meshes, a glass skylight, -ab 1) with the rule on and off are
    equal up to the order of the double-precision atomic additions."""
    octf = golden / "volumes" / "sunroom.oct"
    sens = np.load(golden / "sunroom_sensors.npy")
    out = []
    for off in (False, True):
        if off:
            monkeypatch.setenv("RB_NO_ANYHIT", "1")
        ctx = _lib.Context(0, _lib.RB_PROGRAM_RCONTRIB)
        ctx.load_octree(octf)
        ctx.set_options(["-ab", "1", "-ad", "256", "-lw", "1e-3", "-dc", "1", "-dt", "0", "-dj", "0"])
        ctx.cal_load("reinhart.cal")
        ctx.cal_set("MF=1")
        ctx.add_modifier("solar", "", "rbin", 146)
        out.append((ctx.rcontrib(np.tile(sens, (4, 1)), flags=_lib.RB_IRRAD_RCONTRIB, dtype=np.float64), ctx.stats()["nrays"]))
    monkeypatch.delenv("RB_NO_ANYHIT")
    (a, na), (b, nb) = out
    assert a.sum() > 0 and na <= nb                  # (a blocked ray no longer sends its transmitted child through glass in front)
    np.testing.assert_allclose(a, b, rtol=1e-9, atol=1e-15)


def test_sun_matrix_config5_miniature(G, golden, workdir):
    """BASELINE config 5 in miniature (5-phase direct-sun matrix): 145 `light`
    suns sharing modifier `solar`, reinhart.cal rbin with -e MF:1, louvre
    instances + meshes + glass skylight.  -ab 0 is deterministic: the sun
    coefficients must equal the reference rcontrib's (golden, 1e-5 relative,
    identical zero pattern).  -ab 1 adds the inter-reflected part: row sums
    within 4 % of the oracle's on the flattened scene (64 repetitions pooled)."""
    octf = golden / "volumes" / "sunroom.oct"
    sens = np.load(golden / "sunroom_sensors.npy")
    ref = np.load(golden / "sunroom_ab0.npy")
    rc = pr.Rcontrib(sens.tobytes(), octf, inform="d", outform="d",
                     params=["-I+", "-ab", "0", "-dc", "1", "-dt", "0", "-dj", "0", "-h"])
    rc.add_modifier("solar", calfile="reinhart.cal", expression="MF:1", nbins="Nrbins", binv="rbin")
    m = np.frombuffer(rc(), dtype=np.float64).reshape(48, 146, 3)
    assert np.array_equal(m != 0, ref != 0)
    np.testing.assert_allclose(m, ref, rtol=1e-5, atol=0)
    # inter-reflected sun light, against the oracle on the flattened scene
    ctx = _lib.Context(0, _lib.RB_PROGRAM_RCONTRIB)
    ctx.load_octree(octf)
    flat = workdir / "sunroom_flat_gpu.oct"
    ctx.save_octree(flat)
    ctx.set_options(["-ab", "1", "-ad", "256", "-lw", "1e-3", "-dc", "1", "-dt", "0", "-dj", "0"])
    ctx.cal_load("reinhart.cal")
    ctx.cal_set("MF=1")
    ctx.add_modifier("solar", "", "rbin", 146)
    reps = 8
    big = np.tile(sens, (reps, 1))
    g = ctx.rcontrib(big, flags=_lib.RB_IRRAD_RCONTRIB, dtype=np.float64).reshape(reps, 48, 146, 3).mean(0)
    s = port.Scene(flat, rcontrib=True, ambounce=1, ambdiv=256, minweight=1e-3, dstrsrc=0.0, seed=11)
    s.add_modifier("solar", port.BIN_REINHART, 1, (0, 0, -1), (0, 1, 0), 1.0, 146)
    o = s.rcontrib(big, irrad=2).reshape(reps, 48, 146, 3).mean(0)
    extra_g, extra_o = (g - ref)[:, :, 0].sum(), (o - ref)[:, :, 0].sum()
    assert extra_o > 0.01 * ref[:, :, 0].sum()                  # the bounce adds something measurable
    assert abs(extra_g - extra_o) < 0.04 * extra_o + 1e-6
    # direct part unchanged by the bounce in both
    lit = ref[:, :, 0] > 0
    assert np.all(g[:, :, 0][lit] >= ref[:, :, 0][lit] * (1 - 1e-5))


def test_local_light_sources_vs_reference_golden(golden):
    """local emitters through k_direct -- source partitioning (-ds),
    srcray()'s proximity / spot tests, the aiming test against the source surface,
    shadow rays through glass, m_light on the source.  Deterministic settings
    (-dj 0, every source tested): values within 1e-5 of the reference rtrace /
    rcontrib, identical zero pattern of the coefficient matrix."""
    G = np.load(golden / "lights.npz")
    octf = golden / "lights" / "lights.oct"
    det = ["-ab", "0", "-dt", "0", "-dj", "0", "-dc", "1"]
    for ds in ("0.2", "0", "0.05"):
        ctx = _lib.Context(0)
        ctx.load_octree(octf)
        ctx.set_options(det + ["-ds", ds])
        v, _ = ctx.rtrace(G["sensors"], flags=_lib.RB_IRRAD_RTRACE)
        np.testing.assert_allclose(v, G["irrad_ds" + ds], rtol=1e-5, atol=1e-9)
    ctx = _lib.Context(0)
    ctx.load_octree(octf)
    ctx.set_options(det + ["-ds", ".2"])
    v, _ = ctx.rtrace(G["rays"])
    np.testing.assert_allclose(v, G["view_ds0.2"], rtol=1e-5, atol=1e-9)
    rc = _lib.Context(0, _lib.RB_PROGRAM_RCONTRIB)
    rc.load_octree(octf)
    rc.set_options(["-ab", "0", "-dj", "0", "-ds", ".2"])
    for m in ("lum", "lum2", "spot", "glw", "ill"):
        rc.add_modifier(m, "", "0", 1)
    m = rc.rcontrib(G["sensors"], flags=_lib.RB_IRRAD_RCONTRIB, dtype=np.float64)
    assert np.array_equal(m > 0, G["rcontrib_ab0"] > 0)
    np.testing.assert_allclose(m, G["rcontrib_ab0"], rtol=1e-5, atol=1e-12)
    # jittered, one bounce: totals per emitter against the oracle (8 repetitions pooled, 4 % + noise floor)
    big = np.tile(G["sensors"][:100], (8, 1))
    rc.set_options(["-ab", "1", "-ad", "512", "-lw", "1e-3", "-dj", "0.9", "-ds", ".2"])
    g = rc.rcontrib(big, flags=_lib.RB_IRRAD_RCONTRIB, dtype=np.float64)[:, :, 0].sum(0) / 8
    s = port.Scene(octf, rcontrib=True, ambounce=1, ambdiv=512, minweight=1e-3, srcsizerat=0.2, seed=3)
    for mm in ("lum", "lum2", "spot", "glw", "ill"):
        s.add_modifier(mm)
    o = s.rcontrib(big, irrad=2)[:, :, 0].sum(0) / 8
    assert np.all(np.abs(g - o) <= 0.04 * o + 0.01)


def test_rgbe_output_format(golden):
    """rtrace -fdc -ov and rcontrib -fdc through the Python boundary: the 4-byte
    RGBE stream of the reference, allowing the last mantissa bit to differ where
    a value sits on a rounding boundary (values agree to 1e-6, not bit for bit)."""
    G = np.load(golden / "lights.npz")
    octf = golden / "lights" / "lights.oct"
    out = pr.rtrace(G["rays"].tobytes(), octf, header=False, inform="d", outform="c", outspec="v",
                    params=["-ab", "0", "-dt", "0", "-dj", "0", "-dc", "1", "-ds", ".2"])
    mine = np.frombuffer(out, dtype=np.uint8).reshape(-1, 4).astype(int)
    ref = G["view_rgbe"].astype(int)
    assert mine.shape == ref.shape
    assert np.array_equal(mine[:, 3], ref[:, 3]) and np.abs(mine - ref).max() <= 1 and (mine != ref).mean() < 0.01
    rc = pr.Rcontrib(G["sensors"].tobytes(), octf, inform="d", outform="c", params=["-I", "-ab", "0", "-dj", "0", "-ds", ".2", "-h"])
    for m in G_mods(golden):
        rc.add_modifier(m)
    mine = np.frombuffer(rc(), dtype=np.uint8).reshape(-1, 4).astype(int)
    ref = G["rcontrib_rgbe"].astype(int)
    assert mine.shape == ref.shape
    assert np.abs(mine - ref).max() <= 1 and (mine != ref).mean() < 0.01


def G_mods(golden):
    return json.load(open(golden / "golden.json"))["lights_mods"]


def test_sky_brightness_patterns(golden):
    """brightfunc skies (perezlum.cal `skybright`, skybright.cal `skybr` in its sunny /
    overcast / intermediate branches, with and without a transform) as native device code.  Values seen
    by 2000 rays within 2e-6 of the reference; then `rtrace -I -ab 1 -aa 0` over the reference's own
    trace.oct against the oracle (16 repetitions pooled, 1.5 %)."""
    G = np.load(golden / "sky.npz")
    for name, octf in (("trace", golden / "trace.oct"), ("skies", golden / "sky" / "skies.oct"),
                       ("overcast", golden / "sky" / "overcast.oct")):
        ctx = _lib.Context(0)
        ctx.load_octree(octf)
        ctx.set_options(["-ab", "0"])
        v, _ = ctx.rtrace(G["rays"])
        np.testing.assert_allclose(v, G[name], rtol=2e-6, atol=1e-9)
    sens = np.array([[10, 10, 9.5, 0, 0, 1], [20, 20, 9.5, 0, 0, 1], [10, 10, .5, 0, 0, 1], [3, 3, 2.5, 0, 0, 1]], dtype=float)
    big = np.tile(sens, (16, 1))
    ctx = _lib.Context(0)
    ctx.load_octree(golden / "trace.oct")
    ctx.set_options(["-ab", "1", "-aa", "0", "-as", "0", "-ad", "2048", "-lw", "1e-4", "-dt", "0", "-dj", "0", "-dc", "1"])
    g, _ = ctx.rtrace(big, flags=_lib.RB_IRRAD_RTRACE)
    g = g.reshape(16, 4, 3).mean(0)
    s = port.Scene(golden / "trace.oct", ambounce=1, ambdiv=2048, minweight=1e-4, dstrsrc=0.0, seed=4)
    o = s.rtrace(big, irrad=1)["value"].reshape(16, 4, 3).mean(0)
    np.testing.assert_allclose(g, o, rtol=0.015)
    assert o[2, 0] > 5 and o[0, 0] > 300          # under the ceiling: sky light only through the open sides


def test_rfluxmtx_front_end(golden):
    """rfluxmtx on the CUDA path.  (1) pass-through mode (view rays -> Klems-full window,
    -ab 0) is deterministic: the reference's matrix exactly; (2) sampling mode (Klems window sender ->
    uniform ground + Reinhart MF:2 sky, -ab 0): every sender row sums to 1 and each entry agrees with the
    reference's 5000-sample matrix within 5 sigma of the two binomial estimates (+ 0.003);
    (3) the call of the reference's own test (tests/test_api.py:351-361) in miniature."""
    import os
    F = golden / "flux"
    G = np.load(golden / "flux.npz")
    cwd = os.getcwd()
    os.chdir(F)
    try:
        out = pr.rfluxmtx_main(["rfluxmtx", "-h", "-fdd", "-ab", "0", "-", "window_kf.rad", "room.rad"], G["rays"].tobytes())
        m = np.frombuffer(out, dtype=np.float64).reshape(300, 145, 3)
        assert np.array_equal(m, G["pass_kf"]) and (m[:, :, 0].sum(1) > 0).sum() >= 20
        n = 2000
        out = pr.rfluxmtx_main(["rfluxmtx", "-h", "-ffd", "-ab", "0", "-c", str(n), "sender_window.rad", "sky_r2.rad", "room.rad"],
                               seed=5)
        d = np.frombuffer(out, dtype=np.float64).reshape(145, 578, 3)[:, :, 0]
        ref = G["dmx_kf_r2"].astype(np.float64)
        np.testing.assert_allclose(d.sum(1), 1.0, atol=1e-6)
        pm = 0.5 * (d + ref)
        sig = np.sqrt(pm * (1 - pm) * (1.0 / n + 1.0 / 5000))
        assert np.all(np.abs(d - ref) <= 5 * sig + 0.003)
        assert abs(d[:, 0].sum() - ref[:, 0].sum()) < 1.0          # the ground column as a whole (of 145 rows)
        res = pr.rfluxmtx(F / "sky_r2.rad", rays=b"2 2 1 0 0 1", params=["-ab", "1", "-ad", "64"], scene=[F / "room.rad"])
        assert b"NCOLS=578" in res and b"FORMAT=ascii" in res and len(res.split(b"\n\n", 1)[1].split()) == 578 * 3
        with pytest.raises(RuntimeError, match="-bj"):
            pr.rfluxmtx_main(["rfluxmtx", "-bj", ".5", "-", "window_kf.rad", "room.rad"], b"2 2 1 0 -1 0")
    finally:
        os.chdir(cwd)


def test_matrix_product_on_cta_pairs_matches(monkeypatch):
    """The experimental cta_group::2 form of the tcgen05 matrix product (k_mtx_tc2, RB_MTX_2CTA) gives what the
    default kernel gives, to fp32 rounding, and float64 within the consumer's 1e-5 gate."""
    import ctypes as C
    rng = np.random.default_rng(3)
    nr, ni, nc = 1024, 145, 640
    a = (rng.random((nr, ni, 3)) ** 6 * 0.05).astype(np.float32)
    b = (rng.random((ni, nc, 3)) ** 3 * 2e4).astype(np.float32)
    ref = np.einsum("rik,ick->rck", a.astype(np.float64), b.astype(np.float64))
    outs = []
    for pair in (False, True):
        if pair:
            monkeypatch.setenv("RB_MTX_2CTA", "1")
        ctx = _lib.Context(0)
        out = np.empty((nr, nc, 3), dtype=np.float32)
        ms = C.c_double(0)
        assert ctx.lib.rb_mtx_multiply(ctx.h, a.ctypes.data, nr, ni, b.ctypes.data, nc, out.ctypes.data, 0, C.byref(ms)) == 0
        outs.append(out)
    monkeypatch.delenv("RB_MTX_2CTA")
    for out in outs:
        assert (np.abs(out - ref) / np.maximum(ref, 1e-30)).max() < 1e-5
    np.testing.assert_allclose(outs[0], outs[1], rtol=2e-6)


def test_dctimestep_matrix_consumer(G, golden):
    """the matrix product after the path (rb_mtx_multiply / dctimestep).  Against the reference
    dctimestep on the same files: ascii and float outputs, the -n headerless sky, the V.T.D.s chain -- all
    within 1e-5 (the reference accumulates in double, the kernel in two-level fp32); header lines identical
    except dates; then a large ragged product (K = 2305 like MF:4) against numpy float64."""
    import os
    D = golden / "dct"
    R = np.load(golden / "dct.npz")
    cwd = os.getcwd()
    os.chdir(D)
    try:
        out = pr.dctimestep_main(["dctimestep", "dc.mtx", "sky_f.smx"])
        hdr, body = out.split(b"\n\n", 1)
        ref_hdr = [ln for ln in G["dctimestep_header"].split("\n") if not ln.startswith(("CAPDATE", "GMT"))]
        assert [ln for ln in hdr.decode().split("\n") if not ln.startswith(("CAPDATE", "GMT"))] == ref_hdr
        mine = np.array(body.split(), dtype=np.float64)
        ref = np.array(G["dctimestep_ascii"].split(), dtype=np.float64)
        np.testing.assert_allclose(mine, ref, rtol=1e-5, atol=1e-30)
        assert body.count(b"\n") == 37 and body.split(b"\n")[0].count(b"\t") == 28
        f = np.frombuffer(pr.dctimestep("dc.mtx", "sky_d.smx", header=False, outform="f"), dtype=np.float32).reshape(37, 29, 3)
        np.testing.assert_allclose(f, R["dc_sky"], rtol=1e-5, atol=1e-30)
        assert np.all(f[:, 5] == 0)
        f = np.frombuffer(pr.dctimestep("dc.mtx", (D / "sky_n.txt").read_bytes(), nstep=29, header=False, outform="f"),
                          dtype=np.float32).reshape(37, 29, 3)
        np.testing.assert_allclose(f, R["dc_sky_n"], rtol=1e-5, atol=1e-30)
        f = np.frombuffer(pr.dctimestep("v.mtx", "t.mtx", "d.mtx", "sky_f.smx", header=False, outform="f"),
                          dtype=np.float32).reshape(11, 29, 3)
        np.testing.assert_allclose(f, R["vtds"], rtol=2e-5, atol=1e-30)
        # three-phase form with a Klems BSDF XML as the transmission matrix (util/cmbsdf.c cm_loadBTDF)
        B = np.load(golden / "bsdf.npz")
        f = np.frombuffer(pr.dctimestep("v41.mtx", "bsdf_both.xml", "d41.mtx", "sky_f.smx", header=False, outform="f"),
                          dtype=np.float32).reshape(9, 29, 3)
        np.testing.assert_allclose(f, B["vtds_xml"], rtol=2e-5, atol=1e-30)
        f = np.frombuffer(pr.dctimestep("ident41.mtx", "bsdf_back.xml", "ident41.mtx", "ident41.mtx", header=False, outform="f"),
                          dtype=np.float32).reshape(41, 41, 3)
        assert np.array_equal(f, B["bsdf_back"])                               # identity products are exact
    finally:
        os.chdir(cwd)
    rng = np.random.default_rng(4)
    a = (rng.random((1003, 2305, 3)) ** 5).astype(np.float32)
    b = (rng.random((2305, 517, 3)) ** 2 * 1e4).astype(np.float32)
    ctx = _lib.Context(0)
    c = ctx.mtx_multiply(a, b)
    ref = np.einsum("rik,ick->rck", a.astype(np.float64), b.astype(np.float64))
    np.testing.assert_allclose(c, ref, rtol=1e-5)
    assert ctx.last_mtx_ms > 0


def test_nproc_spreads_records_over_gpus(office2k):
    """`-n N` (Rcontrib(nproc=N)): records go to min(N, visible GPUs) devices inside one process; the
    matrix must be the single-GPU matrix bit for bit (RNG streams are keyed by the global record index)."""
    if _lib.device_count() < 2:
        pytest.skip("needs two visible GPUs (the driver's multi-GPU step / gpurun --gpus 2)")
    sens = scenegen.office_sensors(6000, seed=21)
    args = ["-I+", "-ab", "1", "-ad", "64", "-lw", "1e-2", "-fdf", "-h"] + RB_ARGS
    one = pr.rcontrib_main(["rcontrib", "-n", "1"] + args + [str(office2k)], sens.tobytes())
    two = pr.rcontrib_main(["rcontrib", "-n", "2"] + args + [str(office2k)], sens.tobytes())
    # (the accumulators are doubles filled by atomicAdd: the order of the additions, hence the last bit of a
    #  double sum, is not fixed even on one GPU; float32 rows come out equal except where that bit decides a rounding)
    f32 = lambda b: np.frombuffer(b, dtype=np.float32)
    assert len(one) == 6000 * 145 * 3 * 4
    np.testing.assert_allclose(f32(one), f32(two), rtol=3e-7, atol=0)
    assert (f32(one) != f32(two)).mean() < 1e-3
    eight = pr.rcontrib_main(["rcontrib", "-n", "8", "-c", "3"] + args + [str(office2k)], sens.tobytes())
    one_c = pr.rcontrib_main(["rcontrib", "-c", "3"] + args + [str(office2k)], sens.tobytes())
    assert len(eight) == 2000 * 145 * 3 * 4                              # 2000 records < threshold x GPUs: single path
    np.testing.assert_allclose(f32(eight), f32(one_c), rtol=3e-7, atol=0)
    # rtrace -n N: rays split over the GPUs, random streams keyed by the global ray index
    rays = scenegen.random_rays(140_000, seed=3)
    targs = ["-h", "-fdd", "-ab", "1", "-aa", "0", "-ad", "16", "-lw", "5e-2", "-ovL"]
    t1 = pr.rtrace_main(["rtrace", "-n", "1"] + targs + [str(office2k)], rays.tobytes())
    t2 = pr.rtrace_main(["rtrace", "-n", "2"] + targs + [str(office2k)], rays.tobytes())
    assert len(t1) == 140_000 * 4 * 8
    np.testing.assert_allclose(np.frombuffer(t1), np.frombuffer(t2), rtol=1e-12, atol=0)


_NCCL_WORKER = r"""
import os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.environ["RB_ROOT"])
from pyradiance_b200 import _lib, dist as rbd, scenegen
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", init_method="tcp://127.0.0.1:%s" % os.environ["RB_PORT"], rank=rank, world_size=world,
                        device_id=torch.device("cuda", rank))
octf = os.environ["RB_OCT"]
P = "MF=1,rNx=0,rNy=0,rNz=-1,Ux=0,Uy=1,Uz=0,RHS=+1"
ctx = _lib.Context(rank, _lib.RB_PROGRAM_RCONTRIB)
ctx.load_octree(octf)
ctx.set_options(["-ab", "1", "-ad", "64", "-lw", "1e-2"])
ctx.cal_load("reinhartb.cal"); ctx.cal_set(P)
ctx.add_modifier("skyglow", P, "rbin", 145)
n = 5001                                   # uneven blocks
sens = scenegen.office_sensors(n, seed=21)
flags = _lib.RB_IRRAD_RCONTRIB
single = ctx.rcontrib(sens, flags=flags) if rank == 0 else None
# (1) the peer-memory window: every rank's kernels store their rows into rank 0's HBM
for how in ("window", "host"):
    m = rbd.rcontrib_gathered(ctx, sens, flags=flags, how=how)
    if rank == 0:
        assert m.shape == single.shape
        np.testing.assert_allclose(m, single, rtol=3e-7, atol=0)
        assert (m != single).mean() < 1e-3
    else:
        assert m is None
# (2) blocks kept in each rank's own HBM, exact-size NCCL send / recv into slices of one tensor
mine, r0, r1 = rbd.local_rays(sens, 1, rank, world)
d_rays = torch.from_numpy(mine).cuda()
d_rows = torch.empty((r1 - r0, 145, 3), dtype=torch.float32, device="cuda")
ctx.rcontrib_device(d_rays.data_ptr(), r1 - r0, 1, flags, r0, d_rows.data_ptr(), d_rows.numel())
full = rbd.gather_rows(d_rows, n)
if rank == 0:
    assert full.is_cuda and tuple(full.shape) == (n, 145, 3)
    np.testing.assert_allclose(full.cpu().numpy(), single, rtol=3e-7, atol=0)
    print("DIST_OK")
dist.barrier()
dist.destroy_process_group()
"""


def test_distributed_row_gather_two_gpus(office2k, root, workdir):
    """SURVEY 8e on hardware: one process per GPU (NCCL), records sharded by dist.shard_range, the matrix
    gathered on rank 0 three ways -- peer-memory window (rows stored over NVLink by the finishing kernel),
    shared pinned host matrix (per-GPU D2H), exact-size NCCL send / recv from HBM -- each equal to the
    single-GPU matrix (RNG keyed by the global record index)."""
    if _lib.device_count() < 2:
        pytest.skip("needs two visible GPUs (gpurun --gpus 2)")
    import os
    import socket
    import subprocess
    import sys
    script = workdir / "nccl_worker.py"
    script.write_text(_NCCL_WORKER)
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port_ = s.getsockname()[1]; s.close()
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", RB_ROOT=str(root), RB_PORT=str(port_), RB_OCT=str(office2k))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE))
    outs = [p.communicate(timeout=600) for p in procs]
    assert all(p.returncode == 0 for p in procs), [o[1][-3000:] for o in outs]
    assert b"DIST_OK" in outs[0][0]



def test_smooth_mesh_vertex_normals_vs_reference_golden(golden):
    """mesh triangles with vertex normals (o_mesh.c:193-209 -> RAY.pert ->
    raynormal() in m_normal / m_glass), against the reference rtrace on the same rays
    (tests/golden/make_golden_smooth.py): surface and modifier names, distance, the unperturbed
    (-oN) and perturbed (-on) normals, and the deterministic -ab 0 value, which sees the
    perturbed normal through the sun's cosine and highlight, the mirror direction of the
    metal, and the bent transmission through glass and trans (but not through "Phong")."""
    g = np.load(golden / "smooth.npz")
    rays = g["rays"]
    out = pr.rtrace(rays.tobytes(), str(golden / "smooth" / "smoothroom.oct"), header=False, inform="d", outform="a",
                    outspec="vNnLsm", params=[str(a) for a in g["args"]]).decode()
    rows = [ln.split("\t") for ln in out.splitlines()]
    assert len(rows) == len(rays)
    surf = np.array([r[10] for r in rows]); mod = np.array([r[11] for r in rows])
    assert (surf == g["surf"]).all() and (mod == g["mod"]).all()
    val = np.array([[float(x) for x in r[0:3]] for r in rows])
    fn = np.array([[float(x) for x in r[3:6]] for r in rows])
    pn = np.array([[float(x) for x in r[6:9]] for r in rows])
    dist = np.array([float(r[9]) for r in rows])
    np.testing.assert_allclose(dist, g["dist"], rtol=2e-6)
    np.testing.assert_allclose(fn, g["fnorm"], atol=2e-6)
    np.testing.assert_allclose(pn, g["pnorm"], atol=2e-6)
    smooth = np.abs(np.abs(g["pnorm"]) - np.abs(g["fnorm"])).max(1) > 1e-6
    assert smooth.sum() > 1500
    # values: every ray within tolerance (round 1 allowed two outliers here without naming them; on this build
    # tools/dev_badrows.py lists none)
    bad = ~np.isclose(val, g["value"], rtol=2e-5, atol=1e-7).all(1)
    assert bad.sum() == 0, (bad.sum(), np.flatnonzero(bad)[:10], val[bad][:5], g["value"][bad][:5])
    for m in ("sm_plastic", "sm_metal", "sm_glass", "sm_trans", "Phong", "green"):
        k = (mod == m) & smooth
        assert k.sum() > 30 and (~bad[k]).mean() > 0.97, m


def test_anisotropic_materials_vs_reference_golden(golden):
    """plastic2 / metal2 / trans2 (rt/aniso.c: diraniso, getacoords, agaussamp)
    on the device, against the unmodified reference (tests/golden/make_golden_aniso.py).
    (1) deterministic (-st 1 -dj 0): surface, modifier, distance and value of 2400 view rays
    from above and below the panels, with and without a function transform on the
    orientation vector, and -I values through the trans2 panels: 1e-5 relative;
    (2) highlights sampled (-st 0): per-ray means over 1500 repetitions within 5 combined
    standard errors of the reference's; (3) rcontrib -ab 1 coefficients per tracked emitter:
    means over 200 repetitions within 5 combined standard errors + 2 %."""
    g = np.load(golden / "aniso.npz")
    rays = g["rays"]
    for tag, octf in (("", "aniso.oct"), ("xf_", "anisoxf.oct")):
        out = pr.rtrace(rays.tobytes(), str(golden / "aniso" / octf), header=False, inform="d", outform="a",
                        outspec="vLsm", params=[str(a) for a in g["args"]]).decode()
        rows = [ln.split("\t") for ln in out.splitlines()]
        assert len(rows) == len(rays)
        assert [r[4] for r in rows] == list(g[tag + "surf"]) and [r[5] for r in rows] == list(g[tag + "mod"])
        np.testing.assert_allclose([float(r[3]) for r in rows], g[tag + "dist"], rtol=2e-6)
        val = np.array([[float(x) for x in r[0:3]] for r in rows])
        np.testing.assert_allclose(val, g[tag + "value"], rtol=1e-5, atol=1e-9)
    assert np.abs(g["xf_value"] - g["value"]).max() > 1e-2       # the transform does change the highlights
    ctx = _lib.Context(0)
    ctx.load_octree(golden / "aniso" / "aniso.oct")
    ctx.set_options([str(a) for a in g["args"]])
    v, _ = ctx.rtrace(g["sensors"], flags=_lib.RB_IRRAD_RTRACE)
    np.testing.assert_allclose(v, g["irrad"], rtol=1e-5, atol=1e-9)
    # (2) sampled highlights
    pick, reps = g["st_pick"], 1500
    ctx.set_options([str(a) for a in g["st_args"]])
    v, _ = ctx.rtrace(np.tile(rays[pick], (reps, 1)))
    v = v.reshape(reps, len(pick), 3)
    sem = np.sqrt(v.var(0, ddof=1) / reps + g["st_sem"] ** 2)
    assert (np.abs(v.mean(0) - g["st_mean"]) <= 5 * sem + 1e-5 * g["st_mean"]).all()
    # (3) coefficients per emitter through one diffuse bounce
    rc = _lib.Context(0, _lib.RB_PROGRAM_RCONTRIB)
    rc.load_octree(golden / "aniso" / "aniso.oct")
    rc.set_options([str(a) for a in g["rc_args"]])
    for m in ("skyg", "lampl", "lowl"):
        rc.add_modifier(m, "", "0", 1)
    m = rc.rcontrib(np.tile(rays[pick[:60]], (200, 1)), dtype=np.float64).reshape(200, 60, 3, 3)
    sem = np.sqrt(m.var(0, ddof=1) / 200 + g["rc_sem"] ** 2)
    assert (np.abs(m.mean(0) - g["rc_mean"]) <= 5 * sem + 0.02 * g["rc_mean"] + 1e-9).all()


def test_bsdf_materials_vs_reference_golden(golden, workdir):
    """the BSDF and aBSDF materials (rt/m_bsdf.c over Klems-matrix XML data, common/bsdf_m.c) on the
    device, against the unmodified reference (tests/golden/make_golden_bsdfmat.py): three synthetic BSDF files (Klems
    full with all four blocks, Klems half with one transmission block -> reciprocity, a file-defined basis in Rows
    order, reflection only); aBSDF with its through component, BSDF proxies of both thickness signs over detail
    geometry, an opaque and a thin BSDF, up vectors with and without a function transform, extra diffuse reals.
    (1) nothing sampled (-st 1 -ss 0): surface, modifier, distance and value of 2400 view rays from both sides and -I
    values under / over the panels within 1e-5, with the distant sun only (direct() in the shading thread) and with
    local lamps (k_direct); (2) -ss 1 -st 0: per-ray means of 200 view rays x 1200 repetitions, -I -ab 1 sensors x 150
    and rcontrib coefficients per emitter within 5 combined standard errors of the reference run with -u+;
    (3) what is not built fails by name."""
    g = np.load(golden / "bsdfmat.npz")
    rays = g["rays"]
    D = golden / "bsdfmat"
    for tag, octf in (("", "bsdfmat.oct"), ("lamp_", "bsdflamp.oct")):
        out = pr.rtrace(rays.tobytes(), str(D / octf), header=False, inform="d", outform="a", outspec="vLsm",
                        params=[str(a) for a in g["args"]]).decode()
        rows = [ln.split("\t") for ln in out.splitlines()]
        assert len(rows) == len(rays)
        assert [r[4] for r in rows] == list(g[tag + "surf"]) and [r[5] for r in rows] == list(g[tag + "mod"])
        np.testing.assert_allclose([float(r[3]) for r in rows], g[tag + "dist"], rtol=2e-6)
        val = np.array([[float(x) for x in r[0:3]] for r in rows])
        np.testing.assert_allclose(val, g[tag + "value"], rtol=1e-5, atol=1e-9)
        ctx = _lib.Context(0)
        ctx.load_octree(D / octf)
        ctx.set_options([str(a) for a in g["args"]])
        v, _ = ctx.rtrace(g["sensors"], flags=_lib.RB_IRRAD_RTRACE)
        np.testing.assert_allclose(v, g[tag + "irrad"], rtol=1e-5, atol=1e-9)
    for m in ("awin", "ablind", "prox", "nprox", "opaque", "thin", "avert"):
        assert (g["mod"] == m).sum() >= 80
    # (2) sampled
    pick, reps = g["st_pick"], 1200
    ctx = _lib.Context(0)
    ctx.load_octree(D / "bsdfmat.oct")
    ctx.set_options([str(a) for a in g["st_args"]])
    v, _ = ctx.rtrace(np.tile(rays[pick], (reps, 1)))
    v = v.reshape(reps, len(pick), 3)
    sem = np.sqrt(v.var(0, ddof=1) / reps + g["st_sem"] ** 2)
    assert (np.abs(v.mean(0) - g["st_mean"]) <= 5 * sem + 1e-5 * g["st_mean"]).all()
    assert (v.std(0)[:, 1] > 1e-3 * v.mean(0)[:, 1]).mean() > 0.5         # the samples do vary
    s2, reps2 = g["ab1_sensors"], 150
    ctx = _lib.Context(0)
    ctx.load_octree(D / "bsdfmat.oct")
    ctx.set_options([str(a) for a in g["ab1_args"]])
    v, _ = ctx.rtrace(np.tile(s2, (reps2, 1)), flags=_lib.RB_IRRAD_RTRACE)
    v = v.reshape(reps2, len(s2), 3)
    sem = np.sqrt(v.var(0, ddof=1) / reps2 + g["ab1_sem"] ** 2)
    assert (np.abs(v.mean(0) - g["ab1_mean"]) <= 5 * sem).all()
    rc = _lib.Context(0, _lib.RB_PROGRAM_RCONTRIB)
    rc.load_octree(D / "bsdfmat.oct")
    rc.set_options([str(a) for a in g["rc_args"]])
    for m in ("skyg", "sunl", "gndg"):
        rc.add_modifier(m, "", "0", 1)
    m = rc.rcontrib(np.tile(s2, (reps2, 1)), flags=_lib.RB_IRRAD_RCONTRIB, dtype=np.float64).reshape(reps2, len(s2), 3, 3)
    sem = np.sqrt(m.var(0, ddof=1) / reps2 + g["rc_sem"] ** 2)
    assert (np.abs(m.mean(0) - g["rc_mean"]) <= 5 * sem + 0.004 * g["rc_mean"] + 1e-9).all()
    assert (g["rc_mean"][:, 0] > 0).all() and (g["rc_mean"][:, 1] > 0).any()
    # (3) refusals by name: a tensor-tree file, colour blocks, a missing file, a thickness given as an expression
    xml = (D / "fabric.xml").read_text()
    (workdir / "tt.xml").write_text(xml.replace("<IncidentDataStructure>Columns", "<IncidentDataStructure>TensorTree4"))
    (workdir / "col.xml").write_text(xml.replace(">Visible</Wavelength>", ">CIE-X</Wavelength>", 1))
    for k, (mat, pat) in enumerate(((f"void aBSDF w\n5 {workdir}/tt.xml 0 1 0 .\n0\n0\n", "tensor-tree"),
                                    (f"void aBSDF w\n5 {workdir}/col.xml 0 1 0 .\n0\n0\n", "CIE-X"),
                                    ("void aBSDF w\n5 nosuchfile.xml 0 1 0 .\n0\n0\n", "cannot find BSDF file"),
                                    (f"void BSDF w\n6 thick {D}/fabric.xml 0 1 0 bsdf.cal\n0\n0\n", "unsupported material.*BSDF"))):
        rad = workdir / f"badbsdf{k}.rad"
        rad.write_text(scenegen.MATERIALS + scenegen.SKY + mat + "\nw polygon pane\n0\n0\n12 0 0 1  4 0 1  4 4 1  0 4 1\n\n")
        octf = workdir / f"badbsdf{k}.oct"
        scenegen.build_octree(rad, octf)
        c = _lib.Context(0)
        with pytest.raises(_lib.RBError, match=pat):
            c.load_octree(octf)
            c.set_options(["-ab", "0"])
            c.rtrace(np.array([[2, 2, 3, 0, 0, -1.0]]))


def test_dielectric_interface_vs_reference_golden(golden):
    """dielectric / interface on the device (m_dielectric: Fresnel terms, total
    reflection, refracted direction and solid-angle ratio; the medium id carried by every ray,
    path extinction charged where a ray is shaded, rayorigin()'s weight estimate, shadow rays
    refracted through the bodies) against the unmodified reference
    (tests/golden/make_golden_dielectric.py).  Deterministic (-lr 8): names, distance and value
    of 3000 view rays and the -I values under the bodies within 1e-5; Russian roulette
    (-lr -10 -lw 2e-2): per-ray means over 1200 repetitions within 5 combined standard errors."""
    g = np.load(golden / "dielectric.npz")
    rays = g["rays"]
    octf = golden / "dielectric" / "diel.oct"
    out = pr.rtrace(rays.tobytes(), str(octf), header=False, inform="d", outform="a", outspec="vLsm",
                    params=[str(a) for a in g["args"]]).decode()
    rows = [ln.split("\t") for ln in out.splitlines()]
    assert len(rows) == len(rays)
    assert [r[4] for r in rows] == list(g["surf"]) and [r[5] for r in rows] == list(g["mod"])
    np.testing.assert_allclose([float(r[3]) for r in rows], g["dist"], rtol=2e-6)
    val = np.array([[float(x) for x in r[0:3]] for r in rows])
    # every ray within tolerance (tools/dev_badrows.py lists the ones that are not: none on this build)
    bad = ~np.isclose(val, g["value"], rtol=1e-5, atol=1e-9).all(1)
    assert bad.sum() == 0, (bad.sum(), np.flatnonzero(bad)[:10], val[bad][:5], g["value"][bad][:5])
    ctx = _lib.Context(0)
    ctx.load_octree(octf)
    ctx.set_options([str(a) for a in g["args"]])
    v, _ = ctx.rtrace(g["sensors"], flags=_lib.RB_IRRAD_RTRACE)
    np.testing.assert_allclose(v, g["irrad"], rtol=1e-5, atol=1e-9)
    rc = _lib.Context(0, _lib.RB_PROGRAM_RCONTRIB)           # rcontrib coefficients carry the path extinction (raycontrib)
    rc.load_octree(octf)
    rc.set_options([str(a) for a in g["rc_args"]])
    for m in ("skyg", "lampl", "sunl"):
        rc.add_modifier(m, "", "0", 1)
    cm = rc.rcontrib(rays[:600], dtype=np.float64)
    badc = ~np.isclose(cm, g["rc"], rtol=1e-5, atol=1e-9).reshape(600, -1).all(1)
    assert badc.sum() == 0, (badc.sum(), cm[badc][:3], g["rc"][badc][:3])
    pick, reps = g["rr_pick"], 1200
    ctx.set_options([str(a) for a in g["rr_args"]])
    v, _ = ctx.rtrace(np.tile(rays[pick], (reps, 1)))
    v = v.reshape(reps, len(pick), 3)
    sem = np.sqrt(v.var(0, ddof=1) / reps + g["rr_sem"] ** 2)
    assert (np.abs(v.mean(0) - g["rr_mean"]) <= 5 * sem + 1e-5 * g["rr_mean"]).all()


def test_rmtxop_products_vs_reference(golden, monkeypatch):
    """rmtxop's matrix product `.` on the GPU (rb_mtx_multiply, fp32 with two-level
    accumulation) inside the mirror of the command -- chains, transposes, scalars, component
    transforms before and after -- against the unmodified reference rmtxop (double accumulation,
    float storage): 1e-5 relative to the largest entry of each result, same header and dimensions;
    plus the pyradiance.rmtxop / Rmtxop call forms."""
    from pyradiance_b200 import mtx
    g = np.load(golden / "rmtxop.npz")
    monkeypatch.chdir(golden / "rmtxop")
    n = 0
    for i, c in enumerate(g["cases"]):
        argv = c.split("\x1f")
        ops = [a for a in argv if a in (".", "+", "*", "/")]
        nmat = sum(a.endswith(".mtx") for a in argv)
        if not ("." in ops or nmat - 1 > len(ops)):
            continue
        want = g[f"out{i}"].tobytes()
        got = mtx.rmtxop_main(["rmtxop"] + argv)
        assert got.split(b"\n\n")[0] == want.split(b"\n\n")[0], argv
        a, b = mtx._rmx_parse(got, "got"), mtx._rmx_parse(want, "want")
        assert a.m.shape == b.m.shape and a.dtype == b.dtype
        assert np.abs(a.m - b.m).max() <= 1e-5 * np.abs(b.m).max(), argv
        n += 1
    assert n >= 7
    v = (golden / "rmtxop" / "V.mtx").read_bytes()
    one = mtx.rmtxop(v, outform="f", scale=2.0, transform=[0.265, 0.670, 0.065])
    m = mtx._rmx_parse(one, "one").m
    ref = mtx._rmx_parse(v, "v").m.astype(np.float64) @ np.array([0.265, 0.670, 0.065]) * 2.0
    np.testing.assert_allclose(m[:, :, 0], ref, rtol=1e-6)
    chain = mtx.Rmtxop(outform="d").add_input("V.mtx", transform="Y").add_input("S.mtx", transform="Y")()
    w = mtx._rmx_parse(chain, "chain").m
    assert w.shape == (40, 24, 1)


def test_smooth_mesh_new_materials_vs_reference_golden(golden):
    """RAY.pert through the materials added after the first smooth-mesh fixture: plastic2 / metal2 / trans2
    (perturbed normal in getacoords and diraniso, bent transmission with its guard, the "Phong" exemption)
    and dielectric (perturbed Fresnel angle, the accidental-reflection / penetration guards), on smooth mesh
    triangles placed three ways, against the reference rtrace (tests/golden/make_golden_smooth2.py):
    names, distance, -oN / -on normals and the deterministic value (-st 1)."""
    g0, g = np.load(golden / "smooth.npz"), np.load(golden / "smooth2.npz")
    rays = g0["rays"]
    out = pr.rtrace(rays.tobytes(), str(golden / "smooth" / "smoothroom2.oct"), header=False, inform="d", outform="a",
                    outspec="vNnLsm", params=[str(a) for a in g["args"]]).decode()
    rows = [ln.split("\t") for ln in out.splitlines()]
    assert len(rows) == len(rays)
    surf = np.array([r[10] for r in rows]); mod = np.array([r[11] for r in rows])
    assert (surf == g["surf"]).all() and (mod == g["mod"]).all()
    val = np.array([[float(x) for x in r[0:3]] for r in rows])
    pn = np.array([[float(x) for x in r[6:9]] for r in rows])
    np.testing.assert_allclose([float(r[9]) for r in rows], g["dist"], rtol=2e-6)
    np.testing.assert_allclose(pn, g["pnorm"], atol=2e-6)
    smooth = np.abs(np.abs(g["pnorm"]) - np.abs(g["fnorm"])).max(1) > 1e-6
    bad = ~np.isclose(val, g["value"], rtol=2e-5, atol=1e-7).all(1)
    assert bad.sum() == 0, (bad.sum(), np.flatnonzero(bad)[:10], mod[bad][:10], val[bad][:5], g["value"][bad][:5])
    for m in ("sm_plastic", "sm_metal", "sm_glass", "sm_trans", "Phong"):
        k = (mod == m) & smooth
        assert k.sum() > 100 and (~bad[k]).mean() > 0.98, m
