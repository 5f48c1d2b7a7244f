"""Developer helper (GPU box): rtrace -I -ab 1 means, device vs the live reference, on variants of the BSDF fixture scene."""
import os, re, sys, subprocess
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from pyradiance_b200 import _lib
from oracle import refrun
G = ROOT / "tests" / "golden" / "bsdfmat"
os.environ["RB_RAYPATH_EXTRA"] = str(G)
os.environ["RAYPATH"] = f".:{refrun.LIB}:{G}"
T = Path("/tmp/bsdfvar"); T.mkdir(exist_ok=True)
src = (G / "bsdfmat.rad").read_text()
def variant(name, subs):
    s = src
    for a, b in subs: s = re.sub(a, b, s)
    (T / f"{name}.rad").write_text(s)
    refrun.oconv([T / f"{name}.rad"], T / f"{name}.oct")
    return T / f"{name}.oct"
allplain = [(r"\n(awin|ablind|prox|nprox|opaque|thin|avert) polygon", r"\nplain polygon")]
V = {"full": [], "allplain": allplain,
     "allplain_nosheet": allplain + [(r"\nsheet polygon detail_under_prox\n0\n0\n12[^\n]*\n", "\n"), (r"\nsheet polygon detail_over_nprox\n0\n0\n12[^\n]*\n", "\n")],
     "only_awin": [(r"\n(ablind|prox|nprox|opaque|thin|avert) polygon", r"\nplain polygon")],
     "only_opaque": [(r"\n(ablind|prox|nprox|awin|thin|avert) polygon", r"\nplain polygon")],
     }
sens = np.array([[0.5, 1.5, 3.2, 0, 0, -1], [11.5, .5, 3.2, 0, 0, -1], [10.5, 1.5, .01, 0, 0, 1], [2, 4, 1.2, 0, 1, 0]], float)
args = sys.argv[1:] or ["-ab", "1", "-ad", "512", "-lw", "1e-4", "-dt", "0", "-dj", "0", "-dc", "1", "-st", "0", "-ss", "1", "-aa", "0", "-as", "0"]
reps = 100
for name, subs in V.items():
    octf = variant(name, subs)
    ref = refrun.rtrace(octf, np.tile(sens, (reps, 1)), ["-I"] + args, outform="d").reshape(reps, len(sens), 3)
    ctx = _lib.Context(0); ctx.load_octree(octf); ctx.set_options(args)
    v, _ = ctx.rtrace(np.tile(sens, (reps, 1)), flags=_lib.RB_IRRAD_RTRACE)
    v = v.reshape(reps, len(sens), 3)
    sem = np.sqrt(v.var(0, ddof=1) / reps + ref.var(0, ddof=1) / reps)
    print(name)
    for i in range(len(sens)):
        print("   ", sens[i], "rel dev %", (100 * (v.mean(0)[i] / ref.mean(0)[i] - 1)).round(2), "z", ((v.mean(0)[i] - ref.mean(0)[i]) / sem[i]).round(1))
