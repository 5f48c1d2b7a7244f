"""dctimestep on the GPU: the matrix consumer right after the hot path.

Restates /root/reference/src/radiance/util/dctimestep.c (:163-395 main, matrix forms only) and the
matrix file I/O of util/cmatrix.c (`cm_getheader` :125-166, `cm_load` :191-394, `cm_write` :478-545):
Radiance matrix files with an information header (NROWS= NCOLS= NCOMP=3 [BigEndian=] FORMAT=ascii|float|
double), RGB triplets.  The products run in `rb_mtx_multiply` (csrc/rb_mtx.cu).

Same Python signature as `pyradiance.dctimestep` (src/pyradiance/util.py:140-195).  Not built:
picture inputs (`%03d.hdr` view components) and RGBE output, BSDF XML files as the transmission
matrix, `!command` inputs, per-step output files (`-o spec`).
"""
from __future__ import annotations

import datetime
import os
import re
from pathlib import Path
from typing import Sequence

import numpy as np

from . import _lib
from ._lib import RBError

_FMT = {"ascii": "a", "float": "f", "double": "d"}


def parse_matrix(data: bytes, nrows: int = 0, ncols: int = 0, dtype: str | None = None, what: str = "<stdin>") -> np.ndarray:
    """cm_load(): bytes of a Radiance matrix file -> float32 [nrows, ncols, 3].  `dtype` 'a'/'f'/'d' skips the
    header (like dctimestep -n N / -i?), else the header gives format and dimensions."""
    pos = 0
    swap = False
    scale = np.ones(3)
    dims_ok = (dtype == "a" or nrows > 0) and ncols > 0
    if dtype is None or not dims_ok:
        end = data.find(b"\n\n")
        if not data.startswith(b"#?") or end < 0:
            raise RBError(f"dctimestep: {what}: bad or missing header")
        for line in data[:end].decode("latin-1").split("\n")[1:]:
            if line.startswith("NCOMP="):
                if int(line[6:]) != 3:
                    raise RBError("dctimestep: unexpected # components (must be 3)")
            elif line.startswith("NROWS="):
                nrows = int(line[6:])
            elif line.startswith("NCOLS="):
                ncols = int(line[6:])
            elif line.startswith("BigEndian="):
                swap = line[10:].strip() in ("1", "+", "y", "Y", "t", "T")
            elif line.startswith("EXPOSURE="):
                scale = scale * float(line[9:])
            elif line.startswith("FORMAT="):
                fmt = line[7:].strip()
                if fmt in _FMT:
                    dtype = _FMT[fmt]
                elif "rgbe" in fmt or "xyze" in fmt:
                    raise RBError(f"dctimestep: {what}: picture (RGBE) matrices are not built")
        pos = end + 2
        if dtype is None:
            raise RBError(f"dctimestep: {what}: unexpected data type in header")
    if ncols <= 0:
        raise RBError("dctimestep: unspecified matrix size")
    body = data[pos:]
    if dtype == "a":
        vals = np.array(body.split(), dtype=np.float32)
    else:
        dt = np.dtype(np.float32 if dtype == "f" else np.float64)
        if swap:
            dt = dt.newbyteorder(">")
        vals = np.frombuffer(body, dtype=dt, count=len(body) // dt.itemsize).astype(np.float32)
    per_row = ncols * 3
    if nrows <= 0:
        if vals.size == 0 or vals.size % per_row:
            raise RBError(f"dctimestep: unexpected EOF reading {what}")
        nrows = vals.size // per_row
    if vals.size < nrows * per_row:
        raise RBError(f"dctimestep: unexpected EOF reading {what}")
    m = vals[:nrows * per_row].reshape(nrows, ncols, 3)
    if np.any((scale < .99) | (scale > 1.01)):
        m = (m * scale.astype(np.float32)).astype(np.float32)
    return np.ascontiguousarray(m, dtype=np.float32)


def load_matrix(spec, nrows: int = 0, ncols: int = 0, dtype: str | None = None) -> np.ndarray:
    if isinstance(spec, (bytes, bytearray)):
        return parse_matrix(bytes(spec), nrows, ncols, dtype)
    spec = os.fspath(spec)
    if spec.startswith("!"):
        raise RBError(f"dctimestep: input from command '{spec}' is not supported (commands are not executed)")
    if spec.lower().endswith(".xml"):
        raise RBError("dctimestep: BSDF XML files as the transmission matrix are not built")
    if re.search(r"%[0-9]*[dioxX]", spec):
        raise RBError("dctimestep: picture view components (a %d file specification) are not built")
    try:
        data = Path(spec).read_bytes()
    except OSError:
        raise RBError(f"dctimestep: cannot open file '{spec}'")
    return parse_matrix(data, nrows, ncols, dtype, what=spec)


def multiply(a: np.ndarray, b: np.ndarray, device: int = 0, ctx: _lib.Context | None = None) -> np.ndarray:
    """[nr, ni, 3] x [ni, nc, 3] -> [nr, nc, 3] on the GPU (cm_multiply)."""
    if a.shape[1] <= 0 or a.shape[1] != b.shape[0]:
        raise RBError("dctimestep: matrix dimension mismatch in cm_multiply()")
    own = ctx is None
    ctx = ctx or _lib.Context(device)
    try:
        return ctx.mtx_multiply(a, b)
    finally:
        if own:
            ctx.close()


def _header(argv, nrows, ncols, outfmt) -> bytes:
    now = datetime.datetime.now()
    utc = datetime.datetime.now(datetime.timezone.utc)
    from .rt import _quote_args
    txt = "#?RADIANCE\n" + _quote_args(argv) + "\n"
    txt += now.strftime("CAPDATE= %Y:%m:%d %H:%M:%S\n") + utc.strftime("GMT= %Y:%m:%d %H:%M:%S\n")
    txt += f"NROWS={nrows}\nNCOLS={ncols}\nNCOMP=3\n"
    if outfmt in "fd":
        txt += "BigEndian=0\n"
    txt += "FORMAT=" + {"a": "ascii", "f": "float", "d": "double"}[outfmt] + "\n\n"
    return txt.encode("latin-1")


def dctimestep_main(argv: Sequence[str], stdin: bytes | None = None, device: int = 0) -> bytes:
    """The dctimestep command (argv[0] = program name), matrix forms:
    dctimestep [opts] DCmatrix [sky]   |   dctimestep [opts] Vmatrix Tmatrix Dmatrix [sky]"""
    argv = [str(a) for a in argv]
    skyfmt, outfmt, headout, nsteps, xres, yres = None, "a", True, 0, 0, 0
    a = 1
    usage = RBError("Usage: dctimestep [-n nsteps][-o ospec][-x xr][-y yr][-i{f|d|h}][-o{f|d|c}] DCspec [skyf]\n"
                    "   or: dctimestep [-n nsteps][-o ospec][-x xr][-y yr][-i{f|d|h}][-o{f|d|c}] Vspec Tbsdf Dmat.dat [skyf]")
    while a < len(argv) and argv[a].startswith("-") and len(argv[a]) > 1:
        s = argv[a]
        c = s[1]
        if c == "n":
            nsteps = int(argv[a + 1]); a += 1
            if nsteps < 0:
                raise usage
            skyfmt = "a" if nsteps else None
        elif c == "h":
            headout = not headout
        elif c == "i":
            if s[2:3] not in ("f", "d", "a"):
                raise usage
            skyfmt = s[2]
        elif c == "o":
            if s[2:3] == "":
                raise RBError("dctimestep: per-step output files (-o spec) are not built")
            if s[2] == "c":
                raise RBError("dctimestep: RGBE (-oc) output is not built")
            if s[2] not in "fda":
                raise usage
            outfmt = s[2]
        elif c == "x":
            xres = int(argv[a + 1]); a += 1
        elif c == "y":
            yres = int(argv[a + 1]); a += 1
        else:
            raise usage
        a += 1
    files = argv[a:]
    if not 1 <= len(files) <= 4:
        raise usage

    def sky(spec):
        src = spec if spec is not None else (stdin or b"")
        return load_matrix(src, 0, nsteps, skyfmt)
    ctx = _lib.Context(device)
    try:
        if len(files) > 2:                       # V T D [s]
            smtx = sky(files[3] if len(files) > 3 else None)
            tmat = load_matrix(files[1])
            dmat = load_matrix(files[2], tmat.shape[1], smtx.shape[0])
            cmtx = multiply(tmat, multiply(dmat, smtx, ctx=ctx), ctx=ctx)
        else:
            cmtx = sky(files[1] if len(files) > 1 else None)
        vmat = load_matrix(files[0], 0, cmtx.shape[0])
        r = multiply(vmat, cmtx, ctx=ctx)
    finally:
        ctx.close()
    nr, nc = r.shape[0], r.shape[1]
    if xres > 0 or yres > 0:                     # alt_dim(rmtx, yres, xres)
        n_r, n_c = yres, xres
        if n_r > 0:
            if n_c <= 0:
                n_c = nr * nc // n_r
            if n_r * n_c != nr * nc:
                raise RBError(f"Bad dimensions: {n_r}x{n_c} != {nr}x{nc}")
        else:
            n_r = nr * nc // n_c
            if n_c * n_r != nr * nc:
                raise RBError(f"Bad dimensions: {n_c} does not divide {nr}x{nc} evenly")
        nr, nc = n_r, n_c
        r = r.reshape(nr, nc, 3)
    out = bytearray()
    if headout:
        out += _header(argv, nr, nc, outfmt)
    if outfmt == "a":
        out += _lib.format_ascii(r, triplets=True)
    elif outfmt == "f":
        out += np.ascontiguousarray(r, dtype=np.float32).tobytes()
    else:
        out += np.ascontiguousarray(r, dtype=np.float64).tobytes()
    return bytes(out)


def dctimestep(*mtx, nstep: int | None = None, header: bool = True, xres: int | None = None, yres: int | None = None,
               inform: str | None = None, outform: str | None = None, ospec: str | None = None):
    """Same call as pyradiance.dctimestep (src/pyradiance/util.py:140-195)."""
    cmd = ["dctimestep"]
    if len(mtx) not in (2, 4):
        raise ValueError("mtx must be a list of 2 or 4 items")
    if nstep:
        cmd.extend(["-n", str(nstep)])
    if not header:
        cmd.append("-h")
    if xres:
        cmd.extend(["-x", str(xres)])
    if yres:
        cmd.extend(["-y", str(yres)])
    if inform:
        cmd.append(f"-i{inform}")
    if outform:
        cmd.append(f"-o{outform}")
    if ospec:
        cmd.extend(["-o", ospec])
    stdin = None
    if isinstance(mtx[-1], bytes):
        stdin = mtx[-1]
        mtx = mtx[:-1]
    cmd.extend(os.fspath(m) for m in mtx)
    return dctimestep_main(cmd, stdin)
