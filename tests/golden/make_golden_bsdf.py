"""Regenerates the BSDF-XML fixtures of the dctimestep consumer: synthetic Klems-matrix
BSDF files under tests/golden/dct/ and, from the UNMODIFIED reference dctimestep (oracle/_ref/bin), the
transmission matrices it derives from them (identity view / daylight / sky matrices pick T out exactly)
plus one full V.T.D.s product.  Run in the build container: python tests/golden/make_golden_bsdf.py
"""
import subprocess
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path.insert(0, str(ROOT))
from oracle import refrun  # noqa: E402

D = HERE / "dct"
QUARTER = ("LBNL/Klems Quarter", [0., 9., 27., 45., 63., 90.], [1, 8, 12, 12, 8])
CUSTOM = ("Fixture/Coarse", [0., 12., 40., 70., 90.], [1, 6, 8, 4])        # not one of the built-in bases


def basis_xml(name, tmin, nphis):
    s = f"<AngleBasis><AngleBasisName>{name}</AngleBasisName>\n"
    for i, n in enumerate(nphis):
        s += (f"<AngleBasisBlock><Theta>{(tmin[i] + tmin[i + 1]) / 2 if i else 0}</Theta><nPhis>{n}</nPhis><ThetaBounds>"
              f"<LowerTheta>{tmin[i]}</LowerTheta><UpperTheta>{tmin[i + 1]}</UpperTheta></ThetaBounds></AngleBasisBlock>\n")
    return s + "</AngleBasis>\n"


def block(direction, m, basis):
    body = "\n".join(", ".join("%.6g" % v for v in row) + "," for row in m)
    return (f"<WavelengthData><LayerNumber>System</LayerNumber><Wavelength unit=\"Integral\">Visible</Wavelength>"
            f"<SourceSpectrum>CIE Illuminant D65 1nm.ssp</SourceSpectrum><DetectorSpectrum>ASTM E308 1931 Y.dsp</DetectorSpectrum>"
            f"<WavelengthDataBlock><WavelengthDataDirection>{direction}</WavelengthDataDirection>"
            f"<ColumnAngleBasis>{basis}</ColumnAngleBasis><RowAngleBasis>{basis}</RowAngleBasis>"
            f"<ScatteringDataType>BTDF</ScatteringDataType><ScatteringData>\n{body}\n</ScatteringData></WavelengthDataBlock></WavelengthData>\n")


def bsdf_xml(structure, blocks, basis, junk=""):
    return (junk + '<?xml version="1.0" encoding="UTF-8"?>\n<WindowElement xmlns="http://windows.lbl.gov" '
            'xmlns:xsi="http://www.w3.org/2001/XMLSchema-instance" xsi:schemaLocation="http://windows.lbl.gov BSDF-v1.4.xsd">\n'
            "<WindowElementType>System</WindowElementType>\n<Optical><Layer>\n<Material><Name>fixture</Name><DeviceType>Other</DeviceType></Material>\n"
            f"<DataDefinition><IncidentDataStructure>{structure}</IncidentDataStructure>\n" + basis_xml(*basis) +
            "</DataDefinition>\n" + "".join(blocks) + "</Layer></Optical>\n</WindowElement>\n")


def write_mtx(path, m, fmt="%.6e %.6e %.6e"):
    hdr = f"#?RADIANCE\nNROWS={m.shape[0]}\nNCOLS={m.shape[1]}\nNCOMP=3\nFORMAT=ascii\n\n"
    path.write_text(hdr + "".join("\t".join(fmt % tuple(c) for c in row) + "\n" for row in m))


def ident(n):
    m = np.zeros((n, n, 3))
    m[np.arange(n), np.arange(n)] = 1
    return m


def main():
    rng = np.random.default_rng(21)
    out = {}
    nq, nc = sum(QUARTER[2]), sum(CUSTOM[2])
    # (1) both directions, incident in columns, a diffuse floor well above .01/pi (the loader separates it)
    front = 0.08 + rng.random((nq, nq)) ** 4 * 3.0
    back = 0.02 + rng.random((nq, nq)) ** 4 * 2.0
    (D / "bsdf_both.xml").write_text(bsdf_xml("Columns", [block("Transmission Back", back, QUARTER[0]),
                                                          block("Transmission Front", front, QUARTER[0])], QUARTER))
    # (2) "Transmission Back" only (reciprocity), incident in rows, many zeros and a negative entry, junk
    #     ahead of the XML declaration like genBSDF's recovery line
    sparse = rng.random((nq, nq)) ** 6 * 5.0
    sparse[rng.random((nq, nq)) < .5] = 0
    sparse[3, 7] = -0.25
    (D / "bsdf_back.xml").write_text(bsdf_xml("Rows", [block("Transmission Back", sparse, QUARTER[0])], QUARTER,
                                              junk="Recover using: genBSDF -recover /tmp/x\n"))
    # (3) a basis the file defines itself
    cust = rng.random((nc, nc)) ** 3
    (D / "bsdf_custom.xml").write_text(bsdf_xml("Columns", [block("Transmission Front", cust, CUSTOM[0])], CUSTOM))
    # (4) purely Lambertian transmission: the loader drops the matrix and dctimestep falls back to 145x145
    (D / "bsdf_lamb.xml").write_text(bsdf_xml("Columns", [block("Transmission Front", np.full((nq, nq), 0.11), QUARTER[0])], QUARTER))
    for n in (nq, nc, 145):
        write_mtx(D / f"ident{n}.mtx", ident(n), "%g %g %g")
    env = refrun.environ() if hasattr(refrun, "environ") else None
    for name, n in (("bsdf_both", nq), ("bsdf_back", nq), ("bsdf_custom", nc), ("bsdf_lamb", 145)):
        i = f"ident{n}.mtx"
        r = subprocess.run([str(refrun.BIN / "dctimestep"), "-h", "-of", i, name + ".xml", i, i], cwd=D, env=env, capture_output=True)
        assert r.returncode == 0, r.stderr
        out[name] = np.frombuffer(r.stdout, dtype=np.float32).reshape(n, n, 3)
        print(name, out[name].shape, float(out[name].sum()), r.stderr.decode().strip())
    # a whole three-phase product through bsdf_both.xml
    vm = (rng.random((9, nq, 3)) * 0.2).astype(np.float32)
    dm = (rng.random((nq, 146, 3)) ** 3 * 0.01).astype(np.float32)
    write_mtx(D / "v41.mtx", vm)
    write_mtx(D / "d41.mtx", dm)
    r = subprocess.run([str(refrun.BIN / "dctimestep"), "-h", "-of", "v41.mtx", "bsdf_both.xml", "d41.mtx", "sky_f.smx"],
                       cwd=D, env=env, capture_output=True)
    assert r.returncode == 0, r.stderr
    out["vtds_xml"] = np.frombuffer(r.stdout, dtype=np.float32).reshape(9, 29, 3)
    np.savez_compressed(HERE / "bsdf.npz", **out)
    print("wrote", HERE / "bsdf.npz")


if __name__ == "__main__":
    main()
