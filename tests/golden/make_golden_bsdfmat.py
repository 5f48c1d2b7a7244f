"""Golden fixture for the BSDF / aBSDF materials (rt/m_bsdf.c over Klems-matrix XML data), SURVEY 8f row f4.

TEST INFRASTRUCTURE.  Run in the build container (needs oracle/_ref, the unmodified reference oconv / rtrace /
rcontrib built by oracle/Makefile):

    python tests/golden/make_golden_bsdfmat.py

Writes tests/golden/bsdfmat/{fabric.xml,blind.xml,brdf.xml,bsdfmat.rad,bsdfmat.oct,bsdflamp.rad,bsdflamp.oct} and
tests/golden/bsdfmat.npz.

BSDF files (synthetic, written by this script):
  fabric.xml  Klems full basis, all four blocks: diffuse floor + a forward ("through") peak on the diagonal + a lobe
              around it; reflection with a mirror-direction lobe.
  blind.xml   Klems half basis, "Transmission Front" + both reflections only (rays from the other side go through
              reciprocity), strongly dependent on the exiting azimuth (so the up vector and its transform matter).
  brdf.xml    a basis of its own ("Fixture/Coarse"), IncidentDataStructure Rows, reflection only.

Scene: panels of 1.9 m x 2 m at z = 2 over a grey floor --
  awin     aBSDF fabric.xml                         (through component: view and shadow rays pass, scaled)
  ablind   aBSDF blind.xml, nine reals, function transform -rz 30
  prox     BSDF  fabric.xml, thickness +0.15 with a trans "detail" sheet 5 cm under it (view and shadow rays see the
           detail; ambient / specular rays from above see the BSDF and continue from below the detail)
  nprox    BSDF  blind.xml, thickness -0.12, detail sheet 4 cm above it
  opaque   BSDF  brdf.xml, thickness 0, six reals (reflection only: shadow rays stop)
  thin     BSDF  fabric.xml, thickness 0 -- away from everything at x = 40 (no shadow ray can reach it: the reference's
           direct() is not re-entrant, source.c:419, and a shadow ray that lands on such a surface corrupts it)
  plain    plastic, for reference
and a standing aBSDF window (blind.xml, up = +z) at y = 5; a distant sun, a glow sky above and a glow ground below.
bsdflamp.rad adds two local lamps (then every direct() runs in k_direct on the device).

Deterministic part: -ab 0 -dt 0 -dj 0 -dc 1 -st 1 -ss 0 -av .02 .03 .04 (nothing sampled: the value is the through
component + dir_bsdf() over the sources + the constant ambient term with the unsampled hemispherical scattering
folded in), view rays from above and from below, -I sensors on the floor and under the ceiling.
Stochastic part: -ss 1 -st 0 (BSDF samples and jittered evaluation on): view rays x repetitions, and -I -ab 1
sensors; rcontrib coefficients of the sky through the panels.
"""
import os
import subprocess
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent.parent))
sys.path.insert(0, str(HERE))
from oracle import refrun  # noqa: E402
from make_golden_bsdf import basis_xml, CUSTOM  # noqa: E402

S = HERE / "bsdfmat"
S.mkdir(exist_ok=True)
env = dict(os.environ, RAYPATH=f".:{refrun.LIB}")
os.environ["RB_RAYPATH_EXTRA"] = str(S)          # refrun.* run from another directory: the XML files are found through RAYPATH

FULL = ("LBNL/Klems Full", [0., 5., 15., 25., 35., 45., 55., 65., 75., 90.], [1, 8, 16, 20, 24, 24, 24, 16, 12])
HALF = ("LBNL/Klems Half", [0., 6.5, 19.5, 32.5, 45.5, 58.5, 71.5, 90.], [1, 8, 12, 16, 20, 12, 8])


def patches(basis):
    """centre direction (front exiting), projected solid angle of every patch"""
    _, tmin, nphis = basis
    v, ohm = [], []
    for li, n in enumerate(nphis):
        c0, c1 = np.cos(np.radians(tmin[li])), np.cos(np.radians(tmin[li + 1]))
        pol = 0. if li == 0 else np.radians((tmin[li] + tmin[li + 1]) / 2)
        for k in range(n):
            azi = 2 * np.pi * k / n
            v.append([np.sin(pol) * np.cos(azi), np.sin(pol) * np.sin(azi), np.cos(pol)])
            ohm.append(np.pi * (c0 * c0 - c1 * c1) / n)
    return np.array(v), np.array(ohm)


def block(direction, m, cb, rb):
    body = "\n".join(", ".join("%.6g" % v for v in row) + "," for row in m)
    return (f"<WavelengthData><LayerNumber>System</LayerNumber><Wavelength unit=\"Integral\">Visible</Wavelength>"
            f"<SourceSpectrum>CIE Illuminant D65 1nm.ssp</SourceSpectrum><DetectorSpectrum>ASTM E308 1931 Y.dsp</DetectorSpectrum>"
            f"<WavelengthDataBlock><WavelengthDataDirection>{direction}</WavelengthDataDirection>"
            f"<ColumnAngleBasis>{cb}</ColumnAngleBasis><RowAngleBasis>{rb}</RowAngleBasis>"
            f"<ScatteringDataType>BTDF</ScatteringDataType><ScatteringData>\n{body}\n</ScatteringData></WavelengthDataBlock></WavelengthData>\n")


def bsdf_xml(structure, blocks, bases):
    return ('<?xml version="1.0" encoding="UTF-8"?>\n<!-- synthetic fixture, tests/golden/make_golden_bsdfmat.py -->\n'
            '<WindowElement xmlns="http://windows.lbl.gov" '
            'xmlns:xsi="http://www.w3.org/2001/XMLSchema-instance" xsi:schemaLocation="http://windows.lbl.gov BSDF-v1.4.xsd">\n'
            "<WindowElementType>System</WindowElementType>\n<FileType>BSDF</FileType>\n<Optical><Layer>\n"
            "<Material><Name>fixture</Name><DeviceType>Other</DeviceType></Material>\n"
            f"<DataDefinition><IncidentDataStructure>{structure}</IncidentDataStructure>\n" + "".join(basis_xml(*b) for b in bases) +
            "</DataDefinition>\n" + "".join(blocks) + "</Layer></Optical>\n</WindowElement>\n")


rng = np.random.default_rng(5)
# ---- fabric.xml: m[o][i] (Columns structure: one text row per exiting direction) ----
vf, of = patches(FULL)
n = len(of)
cosang = vf @ vf.T                                  # patch o of the exiting basis against patch i of the incident one:
#   a transmitted "through" ray leaves in patch index o == i (bsdf_m.c fi / bo conventions), its mirror image in the
#   patch rotated by 180 degrees about the normal
lobe = np.exp(-(np.arccos(np.clip(cosang, -1, 1)) / np.radians(14)) ** 2)
lobe /= (lobe * of[:, None]).sum(0, keepdims=True)
T = 0.06 / np.pi + np.diag(0.28 * vf[:, 2] ** 0.5 / of) + 0.22 * lobe * vf[None, :, 2]
T2 = 0.04 / np.pi + np.diag(0.20 * vf[:, 2] ** 0.5 / of) + 0.30 * lobe
mirror = np.array([int(np.argmax(vf @ (v * [-1, -1, 1]))) for v in vf])
R = 0.07 / np.pi + 0.10 * lobe[mirror, :] * (1.2 - vf[None, :, 2])
R2 = 0.12 / np.pi + 0.05 * lobe[mirror, :]
for m in (T, T2, R, R2):
    m *= 1 + 0.02 * rng.standard_normal(m.shape)
(S / "fabric.xml").write_text(bsdf_xml("Columns", [block("Transmission Front", T, FULL[0], FULL[0]),
                                                   block("Transmission Back", T2, FULL[0], FULL[0]),
                                                   block("Reflection Front", R, FULL[0], FULL[0]),
                                                   block("Reflection Back", R2, FULL[0], FULL[0])], [FULL]))
# ---- blind.xml ----
vh, oh = patches(HALF)
lobeh = np.exp(-(np.arccos(np.clip(vh @ vh.T, -1, 1)) / np.radians(20)) ** 2)
lobeh /= (lobeh * oh[:, None]).sum(0, keepdims=True)
down = 0.5 + 0.5 * vh[:, 1]                         # exiting azimuth: more towards +y of the BSDF frame
Tb = 0.03 / np.pi + 0.45 * lobeh * down[:, None] + np.diag(0.10 / oh) * (vh[:, 1] > -0.2)
mirh = np.array([int(np.argmax(vh @ (v * [-1, -1, 1]))) for v in vh])
Rb = 0.05 / np.pi + 0.25 * lobeh[mirh, :] * (1 - 0.8 * down[:, None])
Rb2 = 0.2 / np.pi + 0.02 * lobeh[mirh, :]
(S / "blind.xml").write_text(bsdf_xml("Columns", [block("Transmission Front", Tb, HALF[0], HALF[0]),
                                                  block("Reflection Front", Rb, HALF[0], HALF[0]),
                                                  block("Reflection Back", Rb2, HALF[0], HALF[0])], [HALF]))
# ---- brdf.xml: own basis, Rows ----
vc, oc = patches(CUSTOM)
lobec = np.exp(-(np.arccos(np.clip(vc @ vc.T, -1, 1)) / np.radians(25)) ** 2)
lobec /= (lobec * oc[:, None]).sum(0, keepdims=True)
mirc = np.array([int(np.argmax(vc @ (v * [-1, -1, 1]))) for v in vc])
Rc = 0.02 / np.pi + 0.5 * lobec[mirc, :]
Rc2 = 0.3 * lobec[mirc, :] + 0.001
(S / "brdf.xml").write_text(bsdf_xml("Rows", [block("Reflection Front", Rc.T, CUSTOM[0], CUSTOM[0]),
                                              block("Reflection Back", Rc2.T, CUSTOM[0], CUSTOM[0])], [CUSTOM]))

MATS = """void aBSDF awin
5 fabric.xml 0 1 0 .
0
0

void aBSDF ablind
7 blind.xml 0 1 0 . -rz 30
0
9 .02 .03 .04  .05 .04 .03  .03 .02 .01

void BSDF prox
6 0.15 fabric.xml 0 1 0 .
0
0

void BSDF nprox
6 -0.12 blind.xml 1 0 0 .
0
3 .01 .01 .02

void BSDF opaque
6 0 brdf.xml .3 1 0 .
0
6 .1 .2 .1  .3 .2 .1

void BSDF thin
6 0 fabric.xml 0 1 0 .
0
0

void aBSDF avert
5 blind.xml 0 0 1 .
0
0

void plastic plain
0
0
5 .5 .5 .5 .05 0

void trans sheet
0
0
7 .8 .8 .7 0 0 .6 .9
"""

GEOM = """void plastic grey
0
0
5 .35 .3 .3 0 0

grey polygon floor
0
0
12 -2 -2 0  16 -2 0  16 8 0  -2 8 0

{panels}
sheet polygon detail_under_prox
0
0
12 4 0 1.95  5.9 0 1.95  5.9 2 1.95  4 2 1.95

sheet polygon detail_over_nprox
0
0
12 6 0 2.04  7.9 0 2.04  7.9 2 2.04  6 2 2.04

thin polygon panel_thin
0
0
12 40 0 2  41.9 0 2  41.9 2 2  40 2 2

avert polygon standing
0
0
12 1 5 0.2  9 5 0.2  9 5 3  1 5 3

void light sunl
0
0
3 9000 9000 8000

sunl source sun
0
0
4 .25 -.35 .9 1.5

{lamps}
void glow skyg
0
0
4 .8 .9 1.2 0

skyg source sky
0
0
4 0 0 1 180

void glow gndg
0
0
4 .2 .15 .1 0

gndg source ground
0
0
4 0 0 -1 180
"""

LAMPS = """void light lampl
0
0
3 60 55 40

lampl polygon lamp
0
0
12 3 1 5  9 1 5  9 3 5  3 3 5

void light lowl
0
0
3 30 30 45

lowl polygon lowlamp
0
0
12 2 0 .3  2 2 .3  10 2 .3  10 0 .3

"""

names = ["awin", "ablind", "prox", "nprox", "opaque", "plain"]
panels = ""
for i, m in enumerate(names):
    x0 = 2 * i
    panels += f"{m} polygon panel_{m}\n0\n0\n12 {x0} 0 2  {x0 + 1.9} 0 2  {x0 + 1.9} 2 2  {x0} 2 2\n\n"


def sh(cmd, out=None, stdin=None):
    r = subprocess.run(cmd, cwd=S, env=env, capture_output=True, input=stdin)
    assert r.returncode == 0, r.stderr.decode()
    if r.stderr.strip():
        print("stderr:", r.stderr.decode()[:400])
    if out:
        (S / out).write_bytes(r.stdout)
    return r.stdout


(S / "bsdfmat.rad").write_text(MATS + GEOM.format(panels=panels, lamps=""))
(S / "bsdflamp.rad").write_text(MATS + GEOM.format(panels=panels, lamps=LAMPS))
sh([str(refrun.BIN / "oconv"), "-f", "bsdfmat.rad"], out="bsdfmat.oct")
sh([str(refrun.BIN / "oconv"), "-f", "bsdflamp.rad"], out="bsdflamp.oct")

rng = np.random.default_rng(17)
n = 2400
tgt = np.stack([rng.uniform(-0.5, 12.5, n), rng.uniform(-0.3, 2.3, n), np.full(n, 2.0)], 1)
org = tgt + np.stack([rng.uniform(-3, 3, n), rng.uniform(-3, 3, n), rng.uniform(0.6, 3, n)], 1)
below = rng.random(n) < 0.4
org[below, 2] = rng.uniform(0.3, 1.7, below.sum())
far = rng.random(n) < 0.08                          # the thin BSDF panel, from above and below
tgt[far] = np.stack([rng.uniform(40.1, 41.8, far.sum()), rng.uniform(0.1, 1.9, far.sum()), np.full(far.sum(), 2.0)], 1)
org[far] = tgt[far] + np.stack([rng.uniform(-2, 2, far.sum()), rng.uniform(-2, 2, far.sum()),
                                rng.choice([-1.0, 1.0], far.sum()) * rng.uniform(0.5, 2, far.sum())], 1)
vert = rng.random(n) < 0.1                          # the standing window, from both sides
tgt[vert] = np.stack([rng.uniform(1.2, 8.8, vert.sum()), np.full(vert.sum(), 5.0), rng.uniform(0.4, 2.8, vert.sum())], 1)
org[vert] = tgt[vert] + np.stack([rng.uniform(-2, 2, vert.sum()), rng.choice([-1.0, 1.0], vert.sum()) * rng.uniform(0.5, 2.5, vert.sum()),
                                  rng.uniform(-0.2, 1.5, vert.sum())], 1)
d = tgt - org
d /= np.linalg.norm(d, axis=1, keepdims=True)
rays = np.concatenate([org, d], 1)
det = ["-ab", "0", "-dt", "0", "-dj", "0", "-dc", "1", "-st", "1", "-ss", "0", "-av", ".02", ".03", ".04"]
out = {"rays": rays, "args": np.array(det)}


def trace(octf, rays, args, spec="vLsm"):
    r = sh([str(refrun.BIN / "rtrace"), "-h", "-fda", "-o" + spec] + args + [octf], stdin=rays.tobytes()).decode()
    rows = [ln.split("\t") for ln in r.splitlines()]
    return (np.array([[float(x) for x in q[0:3]] for q in rows]), np.array([float(q[3]) for q in rows]),
            np.array([q[4] for q in rows]), np.array([q[5] for q in rows]))


for tag, octf in (("", "bsdfmat.oct"), ("lamp_", "bsdflamp.oct")):
    v, L, s, m = trace(octf, rays, det)
    out[tag + "value"], out[tag + "dist"], out[tag + "surf"], out[tag + "mod"] = v, L, s, m
    print(tag or "plain", {k: int((m == k).sum()) for k in names + ["thin", "avert", "sheet", "grey"]})

# -I sensors on the floor under the panels (looking up) and above them (looking down), and behind the standing window
sens = np.array([[x, y, z, 0, 0, dz] for x in np.arange(0.5, 12, 1.0) for y in (0.5, 1.5)
                 for z, dz in ((0.01, 1.0), (3.2, -1.0))] +
                [[x, 6.0, 1.2, 0, -1, 0] for x in (2.0, 5.0, 8.0)] + [[x, 4.0, 1.2, 0, 1, 0] for x in (2.0, 5.0, 8.0)], dtype=float)
out["sensors"] = sens
for tag, octf in (("", "bsdfmat.oct"), ("lamp_", "bsdflamp.oct")):
    out[tag + "irrad"] = refrun.rtrace(S / octf, sens, ["-I"] + det, outform="d").reshape(-1, 3)

# The reference's default -u- draws its "random" numbers from a short stratified table indexed by the ray count
# (urand(ilhash(dimlist) + samplendx)): the mean over repetitions of one ray then depends on how many other rays sit
# between them, by a per cent or two.  -u+ (pure Monte Carlo) gives the expectation the device is compared with.
UPLUS = ["-u+"]
# stochastic: BSDF sampling and jittered evaluation on (-ss 1 -st 0), no ambient bounce; mean over repetitions
nst, reps = 200, 1200
pick = np.flatnonzero(np.isin(out["mod"], ["awin", "ablind", "prox", "nprox", "opaque", "thin", "avert", "sheet"]))
pick = pick[np.linspace(0, len(pick) - 1, nst).astype(int)]
st = ["-ab", "0", "-dt", "0", "-dj", "0", "-dc", "1", "-st", "0", "-ss", "1", "-av", ".02", ".03", ".04"]
big = np.tile(rays[pick], (reps, 1))
v = refrun.rtrace(S / "bsdfmat.oct", big, st + UPLUS, outform="d").reshape(reps, len(pick), 3)
out["st_pick"] = pick
out["st_args"] = np.array(st)
out["st_mean"] = v.mean(0)
out["st_sem"] = v.std(0, ddof=1) / np.sqrt(reps)
print("stochastic: mean", out["st_mean"].mean(0), "relative sem (median)",
      np.median(out["st_sem"][:, 1] / np.maximum(out["st_mean"][:, 1], 1e-9)))

# -I -ab 1 on the floor and under the ceiling: ambient rays meet the panels (proxies included)
ab1 = ["-I", "-ab", "1", "-ad", "512", "-lw", "1e-4", "-dt", "0", "-dj", "0", "-dc", "1", "-st", "0", "-ss", "1", "-aa", "0", "-as", "0"]
s2 = sens[::3]
reps2 = 150
v = refrun.rtrace(S / "bsdfmat.oct", np.tile(s2, (reps2, 1)), ab1 + UPLUS, outform="d").reshape(reps2, len(s2), 3)
out["ab1_sensors"], out["ab1_args"] = s2, np.array(ab1[1:])
out["ab1_mean"], out["ab1_sem"] = v.mean(0), v.std(0, ddof=1) / np.sqrt(reps2)
print("ab1: mean", out["ab1_mean"].mean(0), "relative sem (median)", np.median(out["ab1_sem"][:, 1] / np.maximum(out["ab1_mean"][:, 1], 1e-9)))

# rcontrib: what the sky and the sun contribute through one bounce
rc = ["-I", "-ab", "1", "-ad", "512", "-lw", "1e-4", "-dj", "0", "-st", "0", "-ss", "1", "-m", "skyg", "-m", "sunl", "-m", "gndg"]
m = refrun.rcontrib(S / "bsdfmat.oct", np.tile(s2, (reps2, 1)), UPLUS + rc).reshape(reps2, len(s2), 3, 3)
out["rc_args"] = np.array(rc[1:13])
out["rc_mean"], out["rc_sem"] = m.mean(0), m.std(0, ddof=1) / np.sqrt(reps2)
np.savez_compressed(HERE / "bsdfmat.npz", **out)
print("wrote", HERE / "bsdfmat.npz")
