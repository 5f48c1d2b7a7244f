"""Golden digests for the octree builder: sha256 of the octree body (everything after
the information header, whose command line differs) that the UNMODIFIED reference `oconv -f` writes for
seeded synthetic scenes and for the text-scene fixtures of this directory.  The scenes are regenerated
by the tests from the same seeds.  Run in the build container: python tests/golden/make_golden_oct.py
"""
import hashlib
import json
import subprocess
import sys
import tempfile
from pathlib import Path

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path.insert(0, str(ROOT))
from oracle import refrun  # noqa: E402
from pyradiance_b200 import scenegen  # noqa: E402

CASES = {
    "office_3k": dict(npolys=3000, seed=1234),
    "office_20k_2floors": dict(npolys=20000, floors=2, seed=9),
    "office_5k_flat": dict(npolys=5000, seed=5, curved=False),
}
FILES = ["lights/lights.rad", "sky/skies.rad", "flux/room.rad"]


def body_sha(data: bytes) -> str:
    return hashlib.sha256(data[data.index(b"\n\n") + 2:]).hexdigest()


def ref_oconv(rad: Path, extra=()) -> bytes:
    r = subprocess.run([str(refrun.BIN / "oconv"), *extra, "-f", rad.name], cwd=rad.parent, capture_output=True)
    assert r.returncode == 0, r.stderr
    return r.stdout


def main():
    out = {"scenes": {}, "files": {}, "options": {}}
    with tempfile.TemporaryDirectory() as td:
        for name, kw in CASES.items():
            rad = Path(td) / f"{name}.rad"
            scenegen.write_office(rad, **kw)
            out["scenes"][name] = {"args": kw, "sha256": body_sha(ref_oconv(rad))}
        rad = Path(td) / "office_3k.rad"
        out["options"]["office_3k -n 3 -r 2048"] = body_sha(ref_oconv(rad, ["-n", "3", "-r", "2048"]))
    for f in FILES:
        if (HERE / f).exists():
            out["files"][f] = body_sha(ref_oconv(HERE / f))
    (HERE / "octree_sha.json").write_text(json.dumps(out, indent=1) + "\n")
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
