"""Pins the CPU oracle (oracle/rb_oracle.c) against the reference: the golden
vectors of tests/golden/ (made by the unmodified reference binaries, see
tests/golden/make_golden.py) and, when oracle/_ref is present, the reference
rtrace / rcontrib run here on seeded inputs.  No GPU needed."""
import json

import numpy as np
import pytest

from oracle import port, refrun
from pyradiance_b200 import scenegen

BINCASES = [
    ("bins_reinhartb_mf1", port.BIN_REINHARTB, 1, (0, 0, -1), (0, 1, 0)),
    ("bins_reinhartb_mf4", port.BIN_REINHARTB, 4, (0, 0, -1), (0, 1, 0)),
    ("bins_reinhart_mf2", port.BIN_REINHART, 2, (0, 0, 0), (0, 0, 0)),
    ("bins_klems_full", port.BIN_KLEMS_FULL, 1, (0, 0, -1), (0, 1, 0)),
    ("bins_klems_half", port.BIN_KLEMS_HALF, 1, (0, 0, -1), (0, 1, 0)),
    ("bins_klems_quarter", port.BIN_KLEMS_QUARTER, 1, (0, 0, -1), (0, 1, 0)),
    ("bins_hemi", port.BIN_HEMI, 1, (0, 0, -1), (0, 0, 0)),
    ("bins_shirchiu", port.BIN_SHIRCHIU, 6, (0, 0, -1), (0, 1, 0)),
]


@pytest.fixture(scope="module")
def G(golden):
    return json.load(open(golden / "golden.json"))


@pytest.mark.parametrize("name,fn,mf,n,u", BINCASES)
def test_oracle_bins_match_reference(G, golden, name, fn, mf, n, u):
    dirs = np.load(golden / "bin_dirs.npy")[:, 3:]
    want = G[name]["bins"]
    for d, b in zip(dirs, want):
        v = port.bin_of(fn, mf, n, u, 1.0, d)
        got = -1 if v <= -.5 else int(v + .5)
        assert got == b


def test_oracle_known_answer_hits(G, golden):
    k = G["trace_ovposmNL"]
    s = port.Scene(golden / "trace.oct")
    r = s.rtrace(np.array(k["rays"]))
    for i, line in enumerate(k["out"].strip("\n").split("\n")):
        f = line.split("\t")
        assert s.name(r["robj"][i]) == f[9] and s.name(r["omod"][i]) == f[10]
        np.testing.assert_allclose(r["rop"][i], [float(x) for x in f[3:6]], rtol=2e-7, atol=1e-12)
        np.testing.assert_allclose(r["ron"][i], [float(x) for x in f[11:14]], atol=1e-9)
        assert r["rot"][i] == pytest.approx(float(f[14]), rel=2e-7)


def test_oracle_irradiance_ab0(G, golden):
    k = G["trace_I_ab0"]
    s = port.Scene(golden / "trace.oct", dstrsrc=0.0)
    r = s.rtrace(np.array(k["rays"]), irrad=1)
    want = np.array([[float(x) for x in ln.split()] for ln in k["out"].strip().split("\n")])
    np.testing.assert_allclose(r["value"], want, rtol=1e-5, atol=1e-9)     # north_star: -ab 0 within 1e-5


def test_oracle_config1_grid(G, golden):
    gx, gy = np.meshgrid(np.linspace(1, 39, 100), np.linspace(2, 45, 100))
    grid = np.stack([gx.ravel(), gy.ravel(), np.full(10000, 2.5), np.zeros(10000), np.zeros(10000), np.ones(10000)], 1)
    s = port.Scene(golden / "trace.oct", dstrsrc=0.0)
    v = s.rtrace(grid, irrad=1)["value"]
    k = G["trace_grid_I_ab0"]
    assert int((v[:, 0] > 0).sum()) == k["nonzero_rows"]
    assert v.sum() == pytest.approx(k["sum"], rel=1e-6)


@pytest.fixture(scope="module")
def office2k(workdir):
    rad, octf = workdir / "off2k.rad", workdir / "off2k.oct"
    scenegen.write_office(rad, npolys=2000, seed=1)
    scenegen.build_octree(rad, octf)            # own builder (CPU code of librb200.so)
    return octf


@pytest.mark.skipif(not refrun.available(), reason="oracle/_ref not built")
def test_oracle_hits_equal_reference(office2k):
    rays = scenegen.random_rays(20000, seed=11)
    s = port.Scene(office2k)
    r = s.rtrace(rays)
    ref = refrun.rtrace(office2k, rays, ["-ab", "0", "-osmL"]).splitlines()
    assert len(ref) == len(rays)
    for i, line in enumerate(ref):
        f = line.split("\t")
        assert (s.name(r["robj"][i]), s.name(r["omod"][i])) == (f[0], f[1]), i     # bit-exact hit identity
        assert r["rot"][i] == pytest.approx(float(f[2]), rel=1e-6)                # %e prints 7 digits


@pytest.mark.skipif(not refrun.available(), reason="oracle/_ref not built")
def test_oracle_stochastic_agrees_with_reference(golden):
    """Row sums of a -ab 2 coefficient matrix: oracle vs reference, both Monte
    Carlo.  The reference seeds from time(0), so separate runs inside one second
    repeat the same sample; instead ONE reference run carries 16 copies of each
    sensor (consecutive records continue the random sequence -> independent).
    Tolerance: 6 sigma of the difference of the means, sigma estimated from the
    16 repetitions of each arm, plus 0.5 % -- false-alarm rate far below 1e-5."""
    sens = np.array([[10, 10, 3, 0, 0, 1], [4, 5, 3, 0, 0, 1], [20, 20, 12, 0, 0, 1]], dtype=float)
    args = ["-I", "-ab", "2", "-ad", "2048", "-lw", "1e-4", "-f", "reinhartb.cal", "-p",
            "MF=1,rNx=0,rNy=0,rNz=-1,Ux=0,Uy=1,Uz=0,RHS=+1", "-bn", "Nrbins", "-b", "rbin", "-m", "skyglow"]
    reps = 16
    refs = refrun.rcontrib(golden / "contrib.oct", np.tile(sens, (reps, 1)), args)
    refs = refs.reshape(reps, 3, -1, 3)[:, :, :, 0].sum(2)
    mine = []
    for k in range(reps):
        s = port.Scene(golden / "contrib.oct", rcontrib=True, ambounce=2, ambdiv=2048, minweight=1e-4, seed=100 + k)
        s.add_modifier("skyglow", port.BIN_REINHARTB, 1, (0, 0, -1), (0, 1, 0), 1.0, 145)
        mine.append(s.rcontrib(sens, irrad=2)[:, :, 0].sum(1))
    mine = np.array(mine)
    sig = np.sqrt(refs.var(0, ddof=1) / reps + mine.var(0, ddof=1) / reps) + 1e-6
    assert np.all(np.abs(refs.mean(0) - mine.mean(0)) < 6 * sig + 5e-3 * refs.mean(0))
    assert mine[:, 2] == pytest.approx(np.pi, rel=1e-6)        # unobstructed sensor sums to pi


def test_oracle_local_light_sources_vs_reference_golden(golden):
    """local emitters (polygon, triangle, pentagon, sphere, ring,
    cylinder; light / illum / glow-with-radius / spotlight), source partitioning
    (-ds), the aiming test and shadow rays.  Deterministic settings: the oracle
    reproduces the reference's values to float-output precision."""
    G = np.load(golden / "lights.npz")
    octf = golden / "lights" / "lights.oct"
    for ds in ("0.2", "0", "0.05"):
        s = port.Scene(octf, ambounce=0, dstrsrc=0.0, srcsizerat=float(ds))
        v = s.rtrace(G["sensors"], irrad=1)["value"]
        np.testing.assert_allclose(v, G["irrad_ds" + ds], rtol=1e-5, atol=1e-9)
    s = port.Scene(octf, ambounce=0, dstrsrc=0.0, srcsizerat=0.2)
    np.testing.assert_allclose(s.rtrace(G["rays"])["value"], G["view_ds0.2"], rtol=1e-5, atol=1e-9)
    s = port.Scene(octf, rcontrib=True, ambounce=0, dstrsrc=0.0, srcsizerat=0.2)
    for m in ("lum", "lum2", "spot", "glw", "ill"):
        s.add_modifier(m)
    m = s.rcontrib(G["sensors"], irrad=2)
    assert np.array_equal(m > 0, G["rcontrib_ab0"] > 0)
    np.testing.assert_allclose(m, G["rcontrib_ab0"], rtol=1e-5, atol=1e-12)
    assert (G["rcontrib_ab0"][:, :, 0].sum(0) > 0).all()          # every kind of emitter contributes somewhere


def test_oracle_sky_brightness_patterns_vs_reference_golden(golden):
    """brightfunc skies under glow emitters: gen/perezlum.cal (the reference's own
    trace.oct) and gen/skybright.cal overcast + intermediate branches, restated as
    closed forms -- the values the reference rtrace reports for 2000 directions."""
    G = np.load(golden / "sky.npz")
    for name, octf in (("trace", golden / "trace.oct"), ("overcast", golden / "sky" / "overcast.oct")):
        v = port.Scene(octf, ambounce=0).rtrace(G["rays"])["value"]
        np.testing.assert_allclose(v, G[name], rtol=2e-6, atol=1e-9)
    assert G["trace"].min() > 1 and G["overcast"].max() > 10



def test_oracle_anisotropic_materials_vs_reference_golden(golden):
    """plastic2 / metal2 / trans2 (rt/aniso.c) restated in the oracle.
    Deterministic settings (-st 1: highlights folded into the ambient term, -dj 0): every
    view-ray value and the -I values equal the reference's to 1e-5; with the highlights
    sampled (-st 0) the per-ray means over 400 repetitions agree with the reference's
    1500-repetition means within 5 combined standard errors (+ 1e-5 relative)."""
    g = np.load(golden / "aniso.npz")
    octf = golden / "aniso" / "aniso.oct"
    s = port.Scene(octf, ambounce=0, dstrsrc=0.0, specthresh=1.0, ambval=(.02, .03, .04))
    r = s.rtrace(g["rays"])
    assert [s.name(i) for i in r["robj"]] == list(g["surf"])
    np.testing.assert_allclose(r["value"], g["value"], rtol=1e-5, atol=1e-9)
    for m in ("p2x", "m2d", "t2", "p2punt", "m2sharp", "t2diff", "m2ball"):
        assert (g["mod"] == m).sum() > 100
    np.testing.assert_allclose(s.rtrace(g["sensors"], irrad=1)["value"], g["irrad"], rtol=1e-5, atol=1e-9)
    s = port.Scene(octf, ambounce=0, dstrsrc=0.0, specthresh=0.0, ambval=(.02, .03, .04), seed=5)
    pick, reps = g["st_pick"], 400
    v = s.rtrace(np.tile(g["rays"][pick], (reps, 1)))["value"].reshape(reps, len(pick), 3)
    sem = np.sqrt(v.var(0, ddof=1) / reps + g["st_sem"] ** 2)
    assert (np.abs(v.mean(0) - g["st_mean"]) <= 5 * sem + 1e-5 * g["st_mean"]).all()


def test_oracle_bsdf_materials_vs_reference_golden(golden, monkeypatch):
    """BSDF / aBSDF (rt/m_bsdf.c over common/bsdf.c, bsdf_m.c on Klems-matrix XML data) restated in
    the oracle -- its own XML scanner and loader, the BSDF library's queries, m_bsdf() in its recursive form.
    Nothing sampled (-st 1 -ss 0): names, and every view-ray value and -I value of both fixture scenes (sun only /
    with local lamps) equal the reference's to 1e-5; sampled (-st 0 -ss 1): per-ray means over 300 repetitions
    within 5 combined standard errors of the reference's 1200-repetition means (reference run with -u+)."""
    g = np.load(golden / "bsdfmat.npz")
    D = golden / "bsdfmat"
    for tag, octf in (("", "bsdfmat.oct"), ("lamp_", "bsdflamp.oct")):
        s = port.Scene(D / octf, ambounce=0, dstrsrc=0.0, specthresh=1.0, specjitter=0.0, ambval=(.02, .03, .04))
        r = s.rtrace(g["rays"])
        assert [s.name(i) for i in r["robj"]] == list(g[tag + "surf"]) and [s.name(i) for i in r["omod"]] == list(g[tag + "mod"])
        np.testing.assert_allclose(r["value"], g[tag + "value"], rtol=1e-5, atol=1e-9)
        np.testing.assert_allclose(r["rot"], g[tag + "dist"], rtol=2e-6)
        np.testing.assert_allclose(s.rtrace(g["sensors"], irrad=1)["value"], g[tag + "irrad"], rtol=1e-5, atol=1e-9)
    s = port.Scene(D / "bsdfmat.oct", ambounce=0, dstrsrc=0.0, specthresh=0.0, specjitter=1.0, ambval=(.02, .03, .04), seed=9)
    pick, reps = g["st_pick"][::2], 300
    v = s.rtrace(np.tile(g["rays"][pick], (reps, 1)))["value"].reshape(reps, len(pick), 3)
    sem = np.sqrt(v.var(0, ddof=1) / reps + g["st_sem"][::2] ** 2)
    assert (np.abs(v.mean(0) - g["st_mean"][::2]) <= 5 * sem + 1e-5 * g["st_mean"][::2]).all()


def test_oracle_dielectric_interface_vs_reference_golden(golden):
    """dielectric / interface (rt/dielectric.c, no DISPERSE) and the path
    extinction they switch on (rayparticipate, the weight estimate of rayorigin, the distant
    source that an absorbing medium wipes out).  Deterministic settings (-lr 8: no roulette):
    3000 view rays through a tinted slab, a lens and an interface tank, and -I sensors under
    them, equal the reference to 1e-5; with Russian roulette (-lr -10 -lw 2e-2) the per-ray
    means over 300 repetitions agree within 5 combined standard errors."""
    g = np.load(golden / "dielectric.npz")
    octf = golden / "dielectric" / "diel.oct"
    s = port.Scene(octf, ambounce=0, dstrsrc=0.0, specthresh=1.0, ambval=(.05, .05, .05), maxdepth=8, minweight=1e-3)
    r = s.rtrace(g["rays"])
    assert [s.name(i) for i in r["robj"]] == list(g["surf"])
    np.testing.assert_allclose(r["value"], g["value"], rtol=1e-5, atol=1e-9)
    for m in ("tinted", "clear", "water_in_glass"):
        assert (g["mod"] == m).sum() > 500
    np.testing.assert_allclose(s.rtrace(g["sensors"], irrad=1)["value"], g["irrad"], rtol=1e-5, atol=1e-9)
    s = port.Scene(octf, rcontrib=True, ambounce=0, dstrsrc=0.0, specthresh=1.0, maxdepth=8, minweight=1e-3)
    for m in ("skyg", "lampl", "sunl"):
        s.add_modifier(m)
    np.testing.assert_allclose(s.rcontrib(g["rays"][:600]), g["rc"], rtol=1e-5, atol=1e-9)   # coefficients carry the extinction
    s = port.Scene(octf, ambounce=0, dstrsrc=0.0, specthresh=1.0, ambval=(.05, .05, .05), maxdepth=-10, minweight=2e-2, seed=9)
    pick, reps = g["rr_pick"], 300
    v = s.rtrace(np.tile(g["rays"][pick], (reps, 1)))["value"].reshape(reps, len(pick), 3)
    sem = np.sqrt(v.var(0, ddof=1) / reps + g["rr_sem"] ** 2)
    assert (np.abs(v.mean(0) - g["rr_mean"]) <= 5 * sem + 1e-5 * g["rr_mean"]).all()


GEOM_CASES = [("curved", "curved.oct", "curved_rays"), ("coinc", "coinc.oct", "coinc_rays"),
              ("coincfine", "coinc_fine.oct", "coinc_rays")]


@pytest.mark.parametrize("tag,octf,rk", GEOM_CASES)
def test_oracle_cone_family_and_rayreject_vs_reference_golden(golden, tag, octf, rk):
    """every member of the cone family (cone, cup, cylinder, tube, ring),
    sphere and bubble from inside and outside with rays across the end-cap rims, and the tie rules
    of rayreject() (raytrace.c:535-575) on coincident surfaces -- against the unmodified reference
    rtrace (tests/golden/make_golden_geom.py): names exact, distance 1e-9, normal, -ab 0 value 1e-5."""
    g = np.load(golden / "geom.npz")
    s = port.Scene(golden / "geom" / octf, dstrsrc=0.0, ambval=(.1, .1, .1), specthresh=1.0, maxdepth=6, minweight=1e-3)
    r = s.rtrace(g[rk])
    surf = np.array([s.name(i) for i in r["robj"]])
    mod = np.array([s.name(i) for i in r["omod"]])
    assert np.array_equal(surf, g[tag + "_surf"]), np.flatnonzero(surf != g[tag + "_surf"])[:10]
    assert np.array_equal(mod, g[tag + "_mod"])
    loc = g[tag + "_dist"] < 1e9
    np.testing.assert_allclose(r["rot"][loc], g[tag + "_dist"][loc], rtol=1e-9)
    np.testing.assert_allclose(r["ron"][loc], g[tag + "_norm"][loc], atol=1e-9)
    np.testing.assert_allclose(r["value"], g[tag + "_value"], rtol=1e-5, atol=1e-9)


def test_device_walk_restated_on_cpu_equals_reference_walk(golden, office2k, workdir):
    """The CUDA kernel walks the octree with integer cell coordinates, a level-K cell table in place of the
    upper levels, the ray's reciprocal direction in the step and no checked-object set (rb_geom.cuh).
    oracle/rb_oracle.c restates THAT walk in plain C (localhit_dev); here its answers are compared, ray by
    ray and bit for bit, with the recursive restatement of the reference's localhit/raymove/checkhit --
    random rays through the 2k and a 30k-surface office, the cone-family and coincident-surface fixtures
    and the reference's own test octree (rays from outside the scene cube included)."""
    rad, big = workdir / "off30k.rad", workdir / "off30k.oct"
    scenegen.write_office(rad, npolys=30000, seed=9)
    scenegen.build_octree(rad, big)
    g = np.load(golden / "geom.npz")
    cases = [(office2k, scenegen.random_rays(60000, seed=12)), (big, scenegen.random_rays(120000, seed=13)),
             (golden / "geom" / "curved.oct", g["curved_rays"]), (golden / "geom" / "coinc.oct", g["coinc_rays"]),
             (golden / "geom" / "coinc_fine.oct", g["coinc_rays"]),
             (golden / "trace.oct", scenegen.random_rays(40000, seed=5, lo=(-5, -5, -1), hi=(45, 50, 12)))]
    for octf, rays in cases:
        s = port.Scene(octf)
        a = s.rtrace(rays)
        na = s.counters()["nodes"]
        s.reset_counters()
        s.set_walker(1)
        b = s.rtrace(rays)
        assert np.array_equal(a["robj"], b["robj"]) and np.array_equal(a["rot"], b["rot"]), octf
        assert np.array_equal(a["ron"], b["ron"]) and np.array_equal(a["rop"], b["rop"])     # (values draw random numbers)
        assert s.counters()["nodes"] <= na            # the table replaces the upper levels: fewer node words read
