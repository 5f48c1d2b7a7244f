"""Drop-in replacements of pyradiance's process-boundary API for the hot path.

``rtrace()`` and ``Rcontrib`` keep the signatures of the reference
(src/pyradiance/rt.py:278-362 and :165-234).  Like the reference they first
assemble the argv of the ``rtrace`` / ``rcontrib`` program; instead of spawning
the CPU binary, the argv is interpreted here with the option grammar of
rt/rtmain.c:119-349, rt/rcmain.c:207-320 and rt/renderopts.c:123-349 and the
work is done by the CUDA library through the C ABI (include/rb200.h).  Input
bytes and output bytes have the formats of rt/rtrace.c:504-539,907-1009 and
rt/rc2.c:96-125,258-335.  Failures raise RuntimeError, as the reference's
wrappers do for a non-zero exit (src/pyradiance/anci.py:13-30).  There is no
CPU fallback.
"""
from __future__ import annotations

import datetime
import shlex
from pathlib import Path
from typing import Sequence

import numpy as np

from . import _lib
from ._lib import RBError

VERSION_ID = "RADIANCE 6.0a (pyradiance_b200 CUDA sm_100a)"


# --------------------------------------------------------------- helpers ----
def _quote_args(argv):
    """printargs() of common/header.c: words with spaces are quoted."""
    out = []
    for a in argv:
        if a == "" or any(c.isspace() for c in a):
            out.append('"' + a + '"' if '"' not in a else "'" + a + "'")
        else:
            out.append(a)
    return " ".join(out)


def _parse_rays(data: bytes, fmt: str) -> np.ndarray:
    """Input vectors (rt/rtrace.c:504-539 getvec): ascii words, float or double
    triples; origin then direction."""
    if fmt == "a":
        vals = np.array(data.split(), dtype=np.float64) if data.strip() else np.zeros(0)
    elif fmt == "f":
        n = len(data) // 4
        vals = np.frombuffer(data, dtype=np.float32, count=n).astype(np.float64)
    elif fmt == "d":
        n = len(data) // 8
        vals = np.frombuffer(data, dtype=np.float64, count=n)
    else:
        raise RBError(f"botched input format '{fmt}'")
    n6 = (vals.size // 6) * 6
    return np.ascontiguousarray(vals[:n6].reshape(-1, 6))


def _set_format(spec: str, what: str):
    """setformat(): -f[afd][afdc]"""
    if not spec:
        raise RBError(f"{what}: missing format")
    inf = spec[0]
    outf = spec[1] if len(spec) > 1 else spec[0]
    if inf not in "afd" or outf not in "afdc":
        raise RBError(f"{what}: unsupported i/o format '-f{spec}'")
    return inf, outf


def _header(ctx, prog_argv, ncomp, outfmt, extra=""):
    now = datetime.datetime.now()
    utc = datetime.datetime.now(datetime.timezone.utc)
    lines = list(ctx.header_lines())
    lines.append(_quote_args(prog_argv))
    lines.append(f"SOFTWARE= {VERSION_ID}")
    lines.append(now.strftime("CAPDATE= %Y:%m:%d %H:%M:%S"))
    lines.append(utc.strftime("GMT= %Y:%m:%d %H:%M:%S"))
    lines.append(f"NCOMP={ncomp}")
    txt = "\n".join(lines) + "\n" + extra
    if outfmt in "fd":
        txt += "BigEndian=0\n"
    fmt = {"a": "ascii", "f": "float", "d": "double", "c": "32-bit_rle_rgbe"}[outfmt]
    txt += f"FORMAT={fmt}\n\n"
    return txt.encode("latin-1")


def _bool_opt(arg: str, pos: int, cur: bool) -> bool:
    c = arg[pos:pos + 1]
    if c == "":
        return not cur
    if c in "+1":
        return True
    if c in "-0":
        return False
    raise RBError(f"command line error at '{arg}'")


# ---------------------------------------------------------------- rtrace ----
_RT_UNSUPPORTED_SPEC = {"r": "mirrored contribution", "R": "mirrored distance", "x": "unmirrored contribution",
                        "X": "unmirrored distance", "V": "contribution", "W": "coefficient", "l": "effective distance",
                        "t": "ray-tree trace", "T": "source trace"}


def rtrace_main(argv: Sequence[str], stdin: bytes, device: int = 0, _shard=None):
    """Interpret an rtrace command line (argv[0] is the program name)."""
    argv = [str(a) for a in argv]
    if "-version" in argv[1:]:
        return (VERSION_ID + "\n").encode()
    ctx = _lib.Context(device, _lib.RB_PROGRAM_RTRACE)
    try:
        inform, outform = "a", "a"
        outvals = "v"
        header = True
        imm_irrad = False
        lim_dist = False
        hres = vres = 0
        nproc = 1
        i = 1
        while i < len(argv):
            a = argv[i]
            if not a.startswith("-") or len(a) < 2:
                break
            rv = ctx.set_option(argv[i:])
            if rv >= 0:
                i += rv + 1
                continue
            c = a[1]
            if c == "n":
                nproc = max(1, int(argv[i + 1])); i += 1     # number of processes -> number of GPUs
            elif c == "x":
                hres = int(argv[i + 1]); i += 1
            elif c == "y":
                vres = int(argv[i + 1]); i += 1
            elif c == "w":
                pass
            elif c == "I":
                imm_irrad = _bool_opt(a, 2, imm_irrad)
            elif c == "f":
                inform, outform = _set_format(a[2:], "rtrace")
            elif c == "o":
                outvals = a[2:]
            elif c == "h":
                header = _bool_opt(a, 2, header)
            elif c == "l" and a[2:3] == "d":
                lim_dist = _bool_opt(a, 3, lim_dist)
            elif c == "t" and a[2:3] in ("e", "i", "E", "I"):
                pass            # trace include/exclude lists only matter for -ot
            elif c == "P" or c == "p":
                raise RBError(f"unsupported rtrace option '{a}' (persist / primaries)")
            else:
                raise RBError(f"command line error at '{a}'")
            i += 1
        if i != len(argv) - 1:
            raise RBError("missing octree argument" if i >= len(argv) else f"command line error at '{argv[i]}'")
        octree = argv[i]
        if not outvals:
            raise RBError("empty output specification")
        for ch in outvals:
            if ch in _RT_UNSUPPORTED_SPEC:
                raise RBError(f"unsupported output option '-o{ch}' ({_RT_UNSUPPORTED_SPEC[ch]}) in the CUDA path")
            if ch not in "odvLpNnsmMwc~":
                raise RBError(f"unrecognized output option '{ch}'")
        if outform == "c" and outvals != "v":
            raise RBError("color format only with -ov" + (" (-or, -ox are not built)" if outvals[:1] in "rx" else ""))
        want_values = "v" in outvals
        p = ctx.get_params()
        if (imm_irrad or p.do_irrad) and not want_values:
            raise RBError("-I+ and -i+ options require some value output")
        ctx.load_octree(octree)
        rays = _parse_rays(stdin, inform)
        flags = (_lib.RB_IRRAD_RTRACE if imm_irrad else _lib.RB_IRRAD_NONE) | (_lib.RB_FLAG_LIMDIST if lim_dist else 0)
        n = rays.shape[0]
        if vres > 0:
            n = min(n, (hres if hres > 1 else 1) * vres) if hres > 0 or vres > 0 else n
            rays = rays[:n]
        if _shard is not None:                      # a worker of the multi-GPU split below: rays arrive as an array
            return ctx.rtrace(_shard[0], flags=flags, want_values=want_values, want_results=True, row_base=_shard[1])
        ngpu = min(nproc, _lib.device_count()) if nproc > 1 else 1
        if ngpu > 1 and n >= MULTI_GPU_MIN_RAYS * ngpu:
            # -n N: rays go to min(N, visible GPUs) devices, one host thread + context each; the random streams
            # are keyed by the global ray index, so the output is the single-GPU output
            import threading
            bounds = [n * k // ngpu for k in range(ngpu + 1)]
            parts, errors = [None] * ngpu, []

            def work(k):
                try:
                    sl = rays[bounds[k]:bounds[k + 1]]
                    if k == 0:
                        parts[k] = ctx.rtrace(sl, flags=flags, want_values=want_values, want_results=True, row_base=0)
                    else:
                        parts[k] = rtrace_main(argv, b"", device=(device + k) % _lib.device_count(), _shard=(sl, bounds[k]))
                except Exception as e:          # noqa: BLE001 - re-raised in the caller's thread
                    errors.append(e)
            threads = [threading.Thread(target=work, args=(k,)) for k in range(ngpu)]
            for t in threads:
                t.start()
            for t in threads:
                t.join()
            if errors:
                raise errors[0]
            values = np.concatenate([p[0] for p in parts]) if want_values else None
            res = np.concatenate([p[1] for p in parts])
        else:
            values, res = ctx.rtrace(rays, flags=flags, want_values=want_values, want_results=True)
        _report_warnings(ctx, "rtrace")
        out = bytearray()
        ncomp = 0
        for ch in outvals:
            ncomp += {"o": 3, "d": 3, "v": 3, "L": 1, "p": 3, "N": 3, "n": 3, "s": 1, "m": 1, "M": 1, "w": 1,
                      "c": 2, "~": 0}[ch]
        if header:
            out += _header(ctx, ["rtrace"] + argv[1:-1], ncomp, outform)
        if hres > 0 and vres > 0:
            out += f"-Y {vres} +X {hres}\n".encode()
        out += _format_rtrace(ctx, rays, values, res, outvals, outform)
        return bytes(out)
    finally:
        ctx.close()


def _report_warnings(ctx, prog):
    """The reference prints its warnings on stderr ("warning - bad bin number (ignored)", rcontrib.c:306;
    the loader's notes about what it skipped): so does the mirror, once per call."""
    import sys
    try:
        w = ctx.warnings()
        for line in [x for x in w.splitlines() if x.strip()]:
            print(f"{prog}: warning - {line}", file=sys.stderr)
        bad = ctx.stats().get("badbin", 0)
        if bad:
            print(f"{prog}: warning - bad bin number (ignored) for {bad} contributions: -bn is smaller than the bin "
                  "function's range (was -bn given before the -p / -e that sets MF?)", file=sys.stderr)
    except Exception:       # noqa: BLE001 - reporting must never fail a finished run
        pass


def _raynormal(res, rdir):
    """rtrace -on: ron + pert normalised, turned back to the ray's side of the surface when the perturbation
    tipped it over (raytrace.c:445-478), on the ray as the material left it -- m_normal / m_glass reverse a
    surface hit from behind (flipsurface; `pad` = 1 in the result).  pert is zero except on smooth mesh
    triangles (o_mesh.c:201-209)."""
    sgn = np.where(res["pad"] == 1, -1.0, 1.0)
    ron, pert, rod = res["ron"] * sgn[:, None], res["pert"] * sgn[:, None], res["rod"] * sgn
    out = ron.copy()
    sm = (pert * pert).sum(1) > 0
    if sm.any():
        nrm = ron[sm] + pert[sm]
        ln = np.linalg.norm(nrm, axis=1)
        ok = ln > 0
        nrm[ok] /= ln[ok, None]
        nd = -(nrm * rdir[sm]).sum(1)
        fix = ok & ((nd > 0) ^ (rod[sm] > 0))
        nrm[fix] += 2.0 * nd[fix, None] * rdir[sm][fix]
        nrm[~ok] = ron[sm][~ok]
        out[sm] = nrm
    return out


def _names(ctx, idx, none="*", void="void"):
    cache = {}
    out = []
    for i in idx:
        i = int(i)
        if i not in cache:
            cache[i] = ctx.object_name(i) if i >= 0 else None
        out.append(cache[i])
    return out


def _rgbe(rgb: np.ndarray, single: bool = False) -> bytes:
    """Radiance 4-byte RGBE encoding of [n, 3] values.  single=False restates
    setcolr() (common/color.c:797-823, arithmetic in double: what rtrace's -f?c
    does); single=True restates scolor2scolr() (color.c:299-320, the scale factor
    and the products are COLORV = float: what rcontrib's -f?c does)."""
    v = np.ascontiguousarray(rgb, dtype=np.float32 if single else np.float64).reshape(-1, 3)
    d = v.max(axis=1)
    out = np.zeros((v.shape[0], 4), dtype=np.uint8)
    ok = d > 1e-32
    if ok.any():
        m, e = np.frexp(d[ok].astype(np.float64))
        scale = m * 256.0 / d[ok].astype(np.float64)
        if single:
            scale = scale.astype(np.float32)
        prod = v[ok] * scale[:, None]                      # float32 * float32 stays float32
        out[ok, :3] = np.where(v[ok] > 0, prod.astype(np.int64), 0).astype(np.uint8)
        out[ok, 3] = (e + 128).astype(np.uint8)
    return out.tobytes()


def _format_rtrace(ctx, rays, values, res, outvals, outform) -> bytes:
    n = rays.shape[0]
    dirn = rays[:, 3:6]
    norm = np.linalg.norm(dirn, axis=1, keepdims=True)
    bogus = (norm[:, 0] == 0)
    with np.errstate(invalid="ignore", divide="ignore"):
        dnorm = np.where(norm > 0, dirn / np.where(norm > 0, norm, 1), 0.0)
    cols = []           # list of (kind, array/list)
    hit = res["robj"] >= 0
    for ch in outvals:
        if ch == "o":
            cols.append(("r", rays[:, 0:3]))
        elif ch == "d":
            cols.append(("r", dnorm))
        elif ch == "v":
            cols.append(("r", values))
        elif ch == "L":
            cols.append(("r", res["rot"].reshape(-1, 1)))
        elif ch == "p":
            cols.append(("r", res["rop"]))
        elif ch == "N":                # rtrace.c:787-804 oputN: unperturbed normal, flips undone
            cols.append(("r", np.where(hit[:, None], res["ron"], 0.0)))
        elif ch == "n":                # rtrace.c:807-820 oputn: raynormal() of the ray as shading left it
            cols.append(("r", np.where(hit[:, None], _raynormal(res, dnorm), 0.0)))
        elif ch == "w":
            cols.append(("r", res["rweight"].astype(np.float64).reshape(-1, 1)))
        elif ch == "c":
            cols.append(("r", np.zeros((n, 2))))
        elif ch == "s":
            cols.append(("s", [nm if nm is not None else "*" for nm in _names(ctx, res["robj"])]))
        elif ch == "m":
            names = []
            for ro, om in zip(res["robj"], res["omod"]):
                names.append("*" if ro < 0 else (ctx.object_name(int(om)) if om >= 0 else "void"))
            cols.append(("s", names))
        elif ch == "M":
            raise RBError("unsupported output option '-oM' in the CUDA path")
        elif ch == "~":
            cols.append(("s", ["~"] * n))
    if outform == "c":                 # rtrace.c:999-1002: the float radiance through setcolr()
        return _rgbe(values.astype(np.float32).astype(np.float64))
    if outform == "a" and cols and all(kind == "r" for kind, _ in cols):
        return _lib.format_ascii(np.concatenate([np.asarray(c, dtype=np.float64).reshape(n, -1) for _, c in cols], axis=1))
    if outform == "a":
        lines = []
        for i in range(n):
            parts = []
            for kind, c in cols:
                if kind == "r":
                    parts.append("".join("%e\t" % v for v in c[i]))
                else:
                    parts.append(c[i] + "\t")
            lines.append("".join(parts))
        return ("\n".join(lines) + ("\n" if lines else "")).encode("latin-1")
    dt = np.float32 if outform == "f" else np.float64
    if any(kind == "s" for kind, _ in cols):
        out = bytearray()
        for i in range(n):
            for kind, c in cols:
                out += np.asarray(c[i], dtype=dt).tobytes() if kind == "r" else (c[i] + "\t").encode()
        return bytes(out)
    if not cols:
        return b""
    return np.ascontiguousarray(np.concatenate([c for _, c in cols], axis=1), dtype=dt).tobytes()


def rtrace(
    rays: bytes,
    octree,
    header: bool = True,
    inform: str = "a",
    outform: str = "a",
    irradiance: bool = False,
    irradiance_lambertian: bool = False,
    outspec: None | str = None,
    trace_exclude: str = "",
    trace_include: str = "",
    trace_exclude_file=None,
    trace_include_file=None,
    uncorrelated: bool = False,
    xres: None | int = None,
    yres: None | int = None,
    nproc: None | int = None,
    params: None | Sequence[str] = None,
    report: bool = False,
    version: bool = False,
) -> bytes:
    """Run rtrace on the GPU.  Same arguments as pyradiance.rtrace
    (src/pyradiance/rt.py:278-362)."""
    cmd = ["rtrace"]
    if version:
        return rtrace_main(cmd + ["-version"], b"")
    if not isinstance(rays, bytes):
        raise TypeError("Rays must be bytes")
    if not header:
        cmd.append("-h")
    if irradiance:
        cmd.append("-I")
    elif irradiance_lambertian:
        cmd.append("-i")
    cmd.append(f"-f{inform}{outform}")
    if outspec:
        cmd.append(f"-o{outspec}")
    if trace_exclude:
        cmd.append(f"-te{trace_exclude}")
    elif trace_include:
        cmd.append(f"-ti{trace_include}")
    elif trace_exclude_file:
        cmd.append(f"-tE{trace_exclude_file}")
    elif trace_include_file:
        cmd.append(f"-tI{trace_include_file}")
    if uncorrelated:
        cmd.append("-u+")
    if xres is not None:
        cmd.extend(["-x", str(xres)])
    if yres is not None:
        cmd.extend(["-y", str(yres)])
    if nproc:
        cmd.extend(["-n", str(nproc)])
    if params is not None:
        cmd.extend(params)
    cmd.append(str(octree))
    try:
        return rtrace_main(cmd, rays)
    except RBError as e:
        raise RuntimeError(f"rtrace: {e}") from e


# -------------------------------------------------------------- rcontrib ----
def _ofname(ospec: str, mname: str, bn: int):
    """rc2.c:40-92 ofname(): expand %s (modifier) and %d (bin) in an output spec.
    Returns (file name, has_modifier, has_bin)."""
    order = []
    i = 0
    while i < len(ospec):
        if ospec[i] == "%":
            i += 1
            while i < len(ospec) and ospec[i].isdigit():
                i += 1
            c = ospec[i:i + 1]
            if c == "%":
                pass
            elif c == "s":
                if "s" in order:
                    raise RBError(f"bad output format '{ospec}'")
                order.append("s")
            elif c and c in "dioxX":
                if "d" in order:
                    raise RBError(f"bad output format '{ospec}'")
                order.append("d")
            else:
                raise RBError(f"bad output format '{ospec}'")
        i += 1
    args = tuple(mname if o == "s" else bn for o in order)
    return (ospec % args) if args else ospec.replace("%%", "%"), "s" in order, "d" in order


MULTI_GPU_MIN_RECORDS = 2048        # below this a second GPU costs more (octree load) than it saves
MULTI_GPU_MIN_RAYS = 65536           # same for rtrace rays


def rcontrib_main(argv: Sequence[str], stdin: bytes, device: int = 0, return_array: bool = False, _row_base: int = 0,
                  _single: bool = False):
    """Interpret an rcontrib command line (argv[0] is the program name).
    Option order matters exactly as in rt/rcmain.c:207-320: -f/-e/-p act
    immediately, -bn is evaluated when met, -b/-bn/-p/-o are sticky until -m."""
    argv = [str(a) for a in argv]
    if "-version" in argv[1:]:
        return (VERSION_ID + "\n").encode()
    ctx = _lib.Context(device, _lib.RB_PROGRAM_RCONTRIB)
    try:
        inform, outform = "a", "a"
        header = True
        imm_irrad = lim_dist = contrib = force_open = False
        xres = yres = 0
        accumulate = 1
        nproc = 1
        curout = None
        prms, binval, bincnt = "", None, 0
        mods = []                   # (name, outspec, col0, nbins)
        i = 1
        while i < len(argv):
            a = argv[i]
            if not a.startswith("-") or len(a) < 2:
                break
            rv = ctx.set_option(argv[i:])
            if rv >= 0:
                i += rv + 1
                continue

            def need(k=1):
                if i + k >= len(argv):
                    raise RBError(f"command line error at '{a}'")

            def addmod(name):
                col0 = ctx.add_modifier(name, prms, binval if binval is not None else "0", bincnt)
                mods.append((name, curout, col0, ctx.num_columns() - col0))

            c = a[1]
            if c == "n":
                need(); nproc = max(1, int(argv[i + 1])); i += 1
            elif c == "V":
                contrib = _bool_opt(a, 2, contrib)
            elif c == "x":
                need(); xres = int(argv[i + 1]); i += 1
            elif c == "y":
                need(); yres = int(argv[i + 1]); i += 1
            elif c == "w":
                pass
            elif c == "l" and a[2:3] == "d":
                lim_dist = _bool_opt(a, 3, lim_dist)
            elif c == "I":
                imm_irrad = _bool_opt(a, 2, imm_irrad)
            elif c == "f":
                if a[2:3] == "o":
                    force_open = _bool_opt(a, 3, force_open)
                else:
                    inform, outform = _set_format(a[2:], "rcontrib")
            elif c == "o":
                need(); curout = argv[i + 1]; i += 1
            elif c == "r":
                if _bool_opt(a, 2, False):
                    raise RBError("unsupported option: -r (recover) is not built")
            elif c == "h":
                header = _bool_opt(a, 2, header)
            elif c == "p":
                need(); prms = argv[i + 1]; ctx.cal_set(prms); i += 1
            elif c == "c":
                need(); accumulate = int(argv[i + 1]); i += 1
            elif c == "b":
                need()
                if a[2:3] == "n":
                    bincnt = int(ctx.cal_eval(argv[i + 1]) + .5)
                else:
                    binval = argv[i + 1]
                i += 1
            elif c == "m":
                need(); addmod(argv[i + 1]); i += 1
            elif c == "M":
                need()
                for name in Path(argv[i + 1]).read_text().split():
                    addmod(name)
                i += 1
            elif c == "t":
                need(); i += 1
            else:
                raise RBError(f"command line error at '{a}'")
            i += 1
        if not mods:
            raise RBError("missing required modifier argument")
        if i != len(argv) - 1:
            raise RBError("missing octree argument" if i >= len(argv) else f"command line error at '{argv[i]}'")
        ctx.load_octree(argv[i])
        rays = _parse_rays(stdin, inform)
        flags = (_lib.RB_IRRAD_RCONTRIB if imm_irrad else 0) | (_lib.RB_FLAG_LIMDIST if lim_dist else 0) | \
                (_lib.RB_FLAG_CONTRIB if contrib else 0)
        dt = np.float32 if outform in "fc" else np.float64
        nrec = (rays.shape[0] + accumulate - 1) // accumulate if accumulate > 0 else 1
        ngpu = min(nproc, _lib.device_count()) if (nproc > 1 and not _single and accumulate > 0) else 1
        if ngpu > 1 and nrec >= MULTI_GPU_MIN_RECORDS * ngpu:
            # -n N: the reference forks N processes over the records (rc3.c:598-622); here the records go to
            # min(N, visible GPUs) devices, one host thread + context each, rows written straight into
            # disjoint slices of the result.  RNG streams are keyed by the global record index (row_base),
            # so the matrix is the one a single GPU computes.
            import threading
            mat = np.empty((nrec, ctx.num_columns(), 3), dtype=dt)
            bounds = [nrec * k // ngpu for k in range(ngpu + 1)]
            errors = []

            def work(k):
                try:
                    r0, r1 = bounds[k], bounds[k + 1]
                    part = rays[r0 * accumulate:r1 * accumulate]
                    if k == 0:
                        mat[r0:r1] = ctx.rcontrib(part, accum=accumulate, flags=flags, row_base=_row_base + r0, dtype=dt)
                    else:
                        wargv = argv[:-1] + ["-fd" + outform, argv[-1]]
                        mat[r0:r1] = rcontrib_main(wargv, np.ascontiguousarray(part).tobytes(), device=(device + k) % _lib.device_count(),
                                                   return_array=True, _row_base=_row_base + r0, _single=True)
                except Exception as e:          # noqa: BLE001 - re-raised in the caller's thread
                    errors.append(e)
            threads = [threading.Thread(target=work, args=(k,)) for k in range(ngpu)]
            for t in threads:
                t.start()
            for t in threads:
                t.join()
            if errors:
                raise errors[0]
        elif accumulate > 0:
            mat = ctx.rcontrib(rays, accum=accumulate, flags=flags, row_base=_row_base, dtype=dt)
        else:
            # -c 0: one record holding the SUM over all rays (rc2.c:301-302 sf = 1);
            # dummy (zero-direction) rays are ignored (rcontrib.c:392-398)
            live = rays[np.any(rays[:, 3:6] != 0, axis=1)]
            tot = np.zeros((1, ctx.num_columns(), 3), dtype=np.float64)
            chunk = 4096
            for k in range(0, live.shape[0], chunk):
                part = live[k:k + chunk]
                tot += ctx.rcontrib(part, accum=part.shape[0], flags=flags, row_base=k, dtype=np.float64) * part.shape[0]
            mat = tot.astype(dt)
        _report_warnings(ctx, "rcontrib")
        if return_array:
            return mat
        # ---- output streams (rc2.c:150-254 getostream, :339-356 mod_output) ----
        streams = {}                # name (None = stdout) -> dict(cols=[...], info=str)
        for name, ospec, col0, nb in mods:
            for j in range(nb):
                if ospec is None:
                    key, hm, hb = None, False, False
                else:
                    key, hm, hb = _ofname(ospec, name, j)
                st = streams.setdefault(key, {"cols": [], "mod": None, "bin": None})
                if not st["cols"]:
                    st["mod"] = name if hm else None
                    st["bin"] = j if hb else None
                st["cols"].append(col0 + j)
        stdout = bytearray()
        prog_argv = ["rcontrib"] + argv[1:-1]
        for key, st in streams.items():
            reclen = len(st["cols"])
            out = bytearray()
            if header:
                extra = ""
                if key is not None and (st["mod"] is not None or reclen == 1):
                    extra += f"MODIFIER={st['mod'] if st['mod'] is not None else [m for m in mods if m[2] <= st['cols'][0] < m[2] + m[3]][0][0]}\n"
                if key is not None and st["bin"] is not None:
                    extra += f"BIN={st['bin']}\n"
                if yres > 0:
                    extra += f"NROWS={yres * (xres if xres else 1)}\n"
                if xres <= 0 or reclen > 1:
                    extra += f"NCOLS={reclen}\n"
                out += _header(ctx, prog_argv, 3, outform, extra)
            if reclen == 1 and xres > 0 and yres > 0:
                out += f"-Y {yres} +X {xres}\n".encode()
            sub = mat[:, st["cols"], :]
            if outform == "c":         # rc2.c:324-331: float coefficients through scolor_scolr()
                out += _rgbe(sub.reshape(-1, 3), single=True)
            elif outform == "a":
                out += _lib.format_ascii(sub.reshape(sub.shape[0], -1))
            else:
                out += np.ascontiguousarray(sub).tobytes()
            if key is None:
                stdout += out
            elif key.startswith("!"):
                raise RBError("unsupported output spec: pipes to commands ('!cmd') are not built")
            else:
                if Path(key).exists() and not force_open:
                    raise RBError(f"cannot open '{key}' for writing")       # file exists (rc2.c:197-201)
                Path(key).write_bytes(bytes(out))
        return bytes(stdout)
    finally:
        ctx.close()


class Rcontrib:
    """Same construction protocol as pyradiance.Rcontrib (src/pyradiance/rt.py:165-234)."""

    def __init__(self, inp: bytes, octree, nproc: int = 1, yres=None, inform=None, outform=None, report: int = 0,
                 params: None | Sequence[str] = None):
        self.cmd = ["rcontrib"]
        self.octree = octree
        self.inp = inp
        self.cmd.extend(["-n", str(nproc)])
        if params is not None:
            self.cmd.extend(params)
        if None not in (inform, outform):
            self.cmd.append(f"-f{inform}{outform}")
        if yres is not None:
            self.cmd.extend(["-y", str(yres)])
        if report:
            self.cmd.extend(["-t", str(report)])

    def add_modifier(self, modifier=None, modifier_path=None, calfile=None, expression=None, nbins=None, binv=None,
                     param=None, xres=None, yres=None, output=None):
        arglist = []
        if calfile is not None:
            arglist.extend(["-f", str(calfile)])
        if expression is not None:
            arglist.extend(["-e", str(expression)])
        if nbins is not None:
            arglist.extend(["-bn", str(nbins)])
        if binv is not None:
            arglist.extend(["-b", str(binv)])
        if param is not None:
            arglist.extend(["-p", str(param)])
        if xres is not None:
            arglist.extend(["-x", str(xres)])
        if yres is not None:
            arglist.extend(["-y", str(yres)])
        if output is not None:
            arglist.extend(["-o", str(output)])
        if modifier is not None:
            arglist.extend(["-m", modifier])
        elif modifier_path is not None:
            arglist.extend(["-M", modifier_path])
        else:
            raise ValueError("Modifier or modifier path must be provided.")
        self.cmd.extend(arglist)
        return self

    def __call__(self) -> bytes:
        cmd = self.cmd + [str(self.octree)]
        try:
            return rcontrib_main(cmd, self.inp)
        except RBError as e:
            raise RuntimeError(f"rcontrib: {e}") from e

    def as_array(self) -> np.ndarray:
        """Extension: the matrix as float [nrecords, ncols, 3] without the byte round trip."""
        try:
            return rcontrib_main(self.cmd + [str(self.octree)], self.inp, return_array=True)
        except RBError as e:
            raise RuntimeError(f"rcontrib: {e}") from e
