"""In-process API look-alikes of the reference's nanobind module
(src/binding/radiance_ext.cpp) for the hot path: ``RtraceSimulManager``,
``RcontribSimulManager``, ``RayParams`` and the module functions
``get_ray_params / set_ray_params / set_option / initfunc / loadfunc / eval /
set_eparams / calcontext``.

The reference keeps its state in process-global C variables
(rt/RtraceSimulManager.cpp:237,303-308); so does this module: one global
parameter block and one global cal context, which each manager copies into its
own GPU context (rb_ctx) when it loads an octree / prepares output.  Outputs
are caller-owned numpy arrays instead of views into an mmap'd file.  Errors
raise RuntimeError instead of exiting the interpreter.
"""
from __future__ import annotations

import enum

import numpy as np

from . import _lib
from ._lib import RBError

RTdoFIFO, RTtraceSources, RTlimDist, RTimmIrrad, RTmask = 1, 2, 4, 8, 15
RCcontrib = RTmask + 1
RCCONTEXT = "RC."


class RcOutputOp(enum.IntEnum):
    NEW = 0
    FORCE = 1
    RECOVER = 2


# short option names of RayParams (src/binding/params.cpp:182-345) -> rb_params fields
_PFIELDS = {
    "i": "do_irrad", "u": "rand_samp", "dj": "dstrsrc", "dt": "shadthresh", "dc": "shadcert",
    "dr": "directrelay", "dp": "vspretest", "dv": "directvis", "ds": "srcsizerat", "mg": "seccg",
    "ms": "ssampdist", "st": "specthresh", "ss": "specjitter", "bv": "backvis", "lr": "maxdepth",
    "lw": "minweight", "aw": "ambvwt", "aa": "ambacc", "ar": "ambres", "ad": "ambdiv",
    "as_": "ambssamp", "ab": "ambounce",
}
_PVEC = {"me": "cextinction", "ma": "salbedo", "av": "ambval"}


class RayParams:
    """Rendering parameters with the reference's short names (rp.ab, rp.ad, rp.lw, rp.as_ ...)."""

    def __init__(self, raw: _lib.rb_params | None = None):
        object.__setattr__(self, "_p", raw if raw is not None else _lib.rb_params())

    def __getattr__(self, k):
        if k in _PFIELDS:
            v = getattr(self._p, _PFIELDS[k])
            return bool(v) if k in ("i", "u", "bv") else v
        if k in _PVEC:
            return tuple(getattr(self._p, _PVEC[k]))
        raise AttributeError(k)

    def __setattr__(self, k, v):
        if k in _PFIELDS:
            f = _PFIELDS[k]
            cur = getattr(self._p, f)
            setattr(self._p, f, int(v) if isinstance(cur, int) else float(v))
        elif k in _PVEC:
            arr = getattr(self._p, _PVEC[k])
            for i in range(3):
                arr[i] = float(v[i])
        else:
            raise AttributeError(k)

    def copy(self):
        q = _lib.rb_params()
        import ctypes
        ctypes.memmove(ctypes.byref(q), ctypes.byref(self._p), ctypes.sizeof(q))
        return RayParams(q)


class _Globals:
    """The process-global state the reference keeps in C variables."""

    def __init__(self):
        self.params = None          # RayParams (created lazily: needs the library)
        self.cal_ops = []           # replayable cal operations: ("load", f) / ("set", s)
        self.context = ""

    def scratch(self):
        ctx = _lib.Context(0, _lib.RB_PROGRAM_RTRACE)
        return ctx


_G = _Globals()


def _params() -> RayParams:
    if _G.params is None:
        ctx = _G.scratch()
        _G.params = RayParams(ctx.get_params())
        ctx.close()
    return _G.params


def get_ray_params() -> RayParams:
    return _params()


def set_ray_params(rp: RayParams | None = None):
    """ray_restore(): None restores the defaults (rt/raycalls.c:317-377)."""
    if rp is None:
        ctx = _G.scratch()
        _G.params = RayParams(ctx.get_params())
        ctx.close()
    else:
        _G.params = rp.copy() if rp is not _G.params else rp


def set_option(opts):
    """getrenderopt() over a list of words; unknown words are skipped like the
    reference binding does (src/binding/radiance_ext.cpp:134-152)."""
    ctx = _G.scratch()
    ctx.set_params(_params()._p)
    i = 0
    opts = [str(o) for o in opts]
    while i < len(opts):
        rv = ctx.set_option(opts[i:])
        if rv >= 0:
            i += rv
        i += 1
    _G.params = RayParams(ctx.get_params())
    ctx.close()


def ray_done(freall: int = 0):
    return None


def setspectrsamp(cn, wlpt) -> int:
    if list(cn)[:3] != [0, 1, 2] and list(cn)[:3] != [0, 1, 2, 3][:3]:
        raise RuntimeError("unsupported spectral sampling (only RGB is built)")
    return 1


def initfunc():
    _G.cal_ops = []


def calcontext(ctx_name: str):
    _G.context = ctx_name
    return ctx_name


def loadfunc(fname: str):
    c = _G.scratch()
    try:
        c.cal_load(fname)
    except RBError as e:
        raise RuntimeError(str(e)) from e
    finally:
        c.close()
    _G.cal_ops.append(("load", fname))


def set_eparams(params: str):
    _G.cal_ops.append(("set", params))


def _replay_cal(ctx):
    for op, arg in _G.cal_ops:
        if op == "load":
            ctx.cal_load(arg)
        else:
            ctx.cal_set(arg)


def eval(expr: str) -> float:       # noqa: A001 - same name as the reference binding
    c = _G.scratch()
    try:
        _replay_cal(c)
        return c.cal_eval(expr)
    except RBError as e:
        raise RuntimeError(str(e)) from e
    finally:
        c.close()


class Ray:
    """Read-only view of a traced ray (subset of radiance_ext.Ray, :90-123)."""

    def __init__(self, rorg, rdir, res, value, rno):
        self.rorg = tuple(rorg)
        self.rdir = tuple(rdir)
        self.rop = tuple(res["rop"])
        # the callback sees the RAY as shading left it: m_normal / m_glass reverse a surface that was
        # hit from behind (flipsurface, raytrace.c), which `pad` = 1 in the result records
        sgn = -1.0 if int(res["pad"]) == 1 else 1.0
        self.ron = tuple(sgn * x for x in res["ron"])
        self.pert = tuple(sgn * x for x in res["pert"])
        self.rmax = 0.0
        self.rod = sgn * float(res["rod"])
        self.rot = float(res["rot"])
        self.rweight = float(res["rweight"])
        self.rno = rno
        self.rtype = 1
        self.mcol = (0.0, 0.0, 0.0)
        self.rcol = tuple(value)
        self.robj = int(res["robj"])


class RtraceSimulManager:
    """rt/RtraceSimulManager.h:86-177 on the GPU."""

    def __init__(self, octn: str | None = None, device: int = 0):
        self._ctx = _lib.Context(device, _lib.RB_PROGRAM_RTRACE)
        self.rt_flags = 0
        self._cooked = None
        self._trace = None
        self._last_id = 0
        self._loaded = False
        if octn:
            self.load_octree(octn)

    def load_octree(self, octn) -> bool:
        try:
            self._ctx.load_octree(octn)
        except RBError as e:
            raise RuntimeError(str(e)) from e
        self._loaded = True
        return True

    def set_thread_count(self, nt: int = 0) -> int:
        return max(1, nt)

    def ready(self) -> bool:
        return self._loaded

    def set_cooked_call(self, cb):
        self._cooked = cb

    def set_trace_call(self, cb):
        if cb is not None:
            self._trace = cb

    def cleanup_callbacks(self):
        self._cooked = self._trace = None

    def enqueue_bundle(self, orig_direc, rID0: int = 0) -> int:
        if not self._loaded:
            return -1
        od = np.ascontiguousarray(orig_direc, dtype=np.float64).reshape(-1, 6)
        self._ctx.set_params(_params()._p)
        flags = 0
        if self.rt_flags & RTimmIrrad:
            flags |= _lib.RB_IRRAD_MANAGER
        if self.rt_flags & RTlimDist:
            flags |= _lib.RB_FLAG_LIMDIST
        want_v = self._cooked is not None or self._trace is not None
        try:
            vals, res = self._ctx.rtrace(od, flags=flags, want_values=want_v, want_results=True)
        except RBError as e:
            raise RuntimeError(str(e)) from e
        for i in range(od.shape[0]):
            rid = rID0 + i if rID0 else self._last_id + 1
            self._last_id = rid
            ray = Ray(od[i, :3], od[i, 3:], res[i], vals[i] if vals is not None else (0, 0, 0), rid)
            # the GPU path reports the primary ray; the full ray tree is not replayed
            if self._trace is not None:
                self._trace(ray, None)
            if self._cooked is not None:
                if self._cooked(ray, None) is not None and False:
                    return -1
        return od.shape[0]

    def flush_queue(self) -> int:
        return 0

    def cleanup(self, everything: bool = False) -> int:
        if everything:
            self._ctx.close()
            self._loaded = False
        return 0


class RcontribOutput:
    """One output channel (rt/RcontribSimulManager.h:26-110 RcontribOutput): a file that holds a Radiance header,
    `n_rows` rows of `row_bytes` bytes each, and the count of finished rows in its NROWS= field, or -- for an
    empty output specification -- memory only."""
    ROWZERO = "NROWS=0000000000000000\n"           # RcontribSimulManager.cpp:31

    def __init__(self, name, nrows, ncols, dtype=np.float32):
        self._name = name
        self.n_rows = nrows
        self.ncols = ncols                          # coefficient triplets per row
        self.dtype = np.dtype(dtype)
        self.row_bytes = ncols * 3 * self.dtype.itemsize
        self.cur_row = 0
        self.omod = None
        self.obin = -1
        self.beg_data = 0
        self.cols = []                              # (first column in the context's row, count) per modifier
        self._row_count_pos = 0
        self._mm = None
        self.array = None

    def get_name(self):
        return self._name

    # ---- RcontribOutput::NewHeader(), RcontribSimulManager.cpp:452-519 ----
    def _new_header(self, mgr) -> bytes:
        h = "#?RADIANCE\n" + mgr.get_head_str()
        if self.omod:
            h += f"MODIFIER={self.omod}\n"
        if self.obin >= 0:
            h += f"BIN={self.obin}\n"
        self._row_count_pos = len(h) + 6
        h += self.ROWZERO
        esiz = 3 * self.dtype.itemsize
        if (not mgr.xres) or self.row_bytes > esiz:
            h += f"NCOLS={self.row_bytes // esiz}\n"
        h += "NCOMP=3\n"
        h += "BigEndian=0\n"
        h += "FORMAT=" + ("float" if self.dtype == np.float32 else "double")
        align = self.dtype.itemsize
        while (len(h) + 2) % align:                 # data aligned at the end of the header
            h += " "
        h += "\n\n"
        if mgr.xres > 0 and self.row_bytes == esiz:
            h += f"-Y {mgr.yres} +X {mgr.xres}\n"
        self.beg_data = len(h)
        return h.encode("latin-1")

    def open(self, mgr, op):
        """PrepOutput() for this channel; returns the number of rows already there (RECOVER) or 0."""
        done = 0
        if not self._name:                          # memory only
            self.array = np.zeros((max(self.n_rows, 0), self.ncols * 3), dtype=self.dtype)
            return 0
        import os
        path = self._name
        if op == RcOutputOp.RECOVER:
            if not os.path.exists(path):
                raise RuntimeError(f"cannot recover '{path}': no such file")
            done = self._check_header(mgr, path)
        else:
            if op == RcOutputOp.NEW and os.path.exists(path):
                raise RuntimeError(f"cannot open '{path}' for writing (file exists; use RcOutputOp.FORCE or RECOVER)")
            hdr = self._new_header(mgr)
            with open(path, "wb") as f:
                f.write(hdr)
                f.truncate(len(hdr) + self.n_rows * self.row_bytes)
        if self.n_rows > 0:
            self._mm = np.memmap(path, mode="r+", dtype=self.dtype, offset=self.beg_data, shape=(self.n_rows, self.ncols * 3))
            self.array = self._mm
        else:
            self.array = np.zeros((0, self.ncols * 3), dtype=self.dtype)
        return min(done, self.n_rows)

    # ---- RcontribOutput::CheckHeader(), RcontribSimulManager.cpp:537-596 ----
    def _check_header(self, mgr, path) -> int:
        raw = open(path, "rb").read(mgr.get_head_len() + 1024)
        end = raw.find(b"\n\n")
        if end < 0:
            raise RuntimeError(f"cannot find end of header in '{path}'")
        hdr = raw[:end + 1].decode("latin-1")

        def arg(key):
            for line in hdr.split("\n"):
                if line.startswith(key):
                    return line[len(key):]
            return None
        if int(arg("NCOMP=") or 3) != 3:
            raise RuntimeError(f"expected NCOMP=3 in '{path}'")
        fmt = "float" if self.dtype == np.float32 else "double"
        if not (arg("FORMAT=") or "").startswith(fmt):
            raise RuntimeError(f"expected FORMAT={fmt} in '{path}'")
        esiz = 3 * self.dtype.itemsize
        nc = arg("NCOLS=")
        if (int(nc) * esiz != self.row_bytes) if nc is not None else ((not mgr.xres) or self.row_bytes > esiz):
            raise RuntimeError(f"expected NCOLS={self.row_bytes // esiz} in '{path}'")
        nr = arg("NROWS=")
        if nr is None:
            raise RuntimeError(f"missing NROWS in '{path}'")
        self._row_count_pos = hdr.index("NROWS=") + 6
        self.beg_data = end + 2
        if mgr.xres > 0 and self.row_bytes == esiz:
            res = f"-Y {mgr.yres} +X {mgr.xres}\n".encode()
            if raw[self.beg_data:self.beg_data + len(res)] != res:
                raise RuntimeError(f"bad resolution string in '{path}'")
            self.beg_data += len(res)
        return int(nr)

    def rows_done(self, n):
        """Record the number of finished rows in the header (the file is valid at any time)."""
        self.cur_row = n
        if self._mm is not None:
            self._mm.flush()
            with open(self._name, "r+b") as f:
                f.seek(self._row_count_pos)
                f.write(b"%016d" % n)

    def close(self):
        if self._mm is not None:
            self._mm.flush()
            self._mm = None
        self.array = None


class RcontribSimulManager:
    """rt/RcontribSimulManager.h:173-355 on the GPU."""

    def __init__(self, octn: str | None = None, device: int = 0):
        self._ctx = _lib.Context(device, _lib.RB_PROGRAM_RCONTRIB)
        self._flags = 0
        self.accum = 1
        self.xres = 0
        self.yres = 0
        self.out_op = RcOutputOp.NEW
        self.cds_f = None
        self._dtype = np.float32
        self._mods = []             # (modn, outspec, prms, binval, bincnt, cal_ops snapshot)
        self._loaded = False
        self._out = None
        self._outputs = []          # RcontribOutput per distinct output specification
        self._rows_done = 0
        self._device = device
        self._header = ""
        if octn:
            self.load_octree(octn)

    # ---- header (rt/RtraceSimulManager.cpp:46-171 RadSimulManager::NewHeader / AddHeader / GetHeadStr) ----
    def new_header(self, inspec=None) -> bool:
        """Prepare the header from a previous input (a file with a Radiance header, or `!command`), or clear it."""
        self._header = ""
        if not inspec:
            return False
        if inspec[0] == "!":
            return self.add_header(inspec[1:])
        try:
            with open(inspec, "rb") as f:
                for raw in f:
                    line = raw.decode("latin-1")
                    if line in ("\n", "\r\n"):
                        break
                    if line.startswith("#?") or line.startswith("FORMAT="):
                        continue
                    self.add_header(line)
        except OSError:
            return False
        return True

    def add_header(self, s) -> bool:
        """Add a line (a string; a newline is added if missing) or a program line (a list of arguments, quoted
        where they hold blanks or quotes) to the header."""
        if s is None:
            return False
        if not isinstance(s, str):
            words = []
            for a in s:
                a = str(a)
                if a == "" or any(c.isspace() for c in a) or '"' in a or "'" in a:
                    qc = "'" if '"' in a else '"'
                    words.append(qc + a + qc)
                else:
                    words.append(a)
            if not words:
                return False
            self._header += " ".join(words) + "\n"
            return True
        s = s.rstrip("\r\n")
        if not s:
            return False
        self._header += s + "\n"
        return True

    def get_head_len(self) -> int:
        return len(self._header.encode("latin-1"))

    def get_head_str(self, key=None, inOK: bool = False):
        """The header lines, or -- with `key` -- what follows the first line that starts with it (None if absent)."""
        if key is None:
            return self._header
        if not key or not self._header or "\n" in key:
            return None
        if inOK:
            key = key.lstrip()
        if not key:
            return None
        for line in self._header.split("\n"):
            probe = line.lstrip() if inOK else line
            if probe.startswith(key):
                return probe[len(key):]
        return None

    def get_format(self, siz=None) -> int:
        """Current data format as the reference's character code ('f' or 'd')."""
        return ord("f") if self._dtype == np.float32 else ord("d")

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.cleanup(True)
        return False

    def has_flag(self, fl) -> bool:
        return bool(self._flags & fl)

    def set_flag(self, fl, val: bool = True) -> bool:
        self._flags = (self._flags | fl) if val else (self._flags & ~fl)
        return True

    def set_data_format(self, ty) -> bool:
        ch = chr(ty) if isinstance(ty, int) else str(ty)
        if ch == "f":
            self._dtype = np.float32
        elif ch == "d":
            self._dtype = np.float64
        else:
            raise RuntimeError(f"unsupported data format '{ch}' (RGBE output is not built)")
        return True

    def load_octree(self, octn) -> bool:
        try:
            self._ctx.load_octree(octn)
        except RBError as e:
            raise RuntimeError(str(e)) from e
        self._loaded = True
        self._header = ""               # NewHeader(octn): the octree's header without its id and FORMAT lines
        for line in self._ctx.header_lines():
            if not (line.startswith("#?") or line.startswith("FORMAT=")):
                self.add_header(line)
        return True

    def add_modifier(self, modn, outspec, prms="", binval="", bincnt=1) -> bool:
        if outspec and "%d" in outspec:
            raise RuntimeError("unsupported output specification: one file per bin ('%d') is not built in this manager")
        try:
            # bin functions come from the global cal context at the time of the call
            scratch = _lib.Context(self._device, _lib.RB_PROGRAM_RCONTRIB)
            try:
                _replay_cal(scratch)
                scratch.add_modifier(modn, prms or "", binval or "0", int(bincnt))
            finally:
                scratch.close()
        except RBError as e:
            raise RuntimeError(str(e)) from e
        self._mods.append((modn, outspec, prms or "", binval or "0", int(bincnt), list(_G.cal_ops)))
        return True

    def add_mod_file(self, modfn, outspec, prms=None, binval=None, bincnt=1) -> bool:
        with open(modfn) as f:
            for name in f.read().split():
                self.add_modifier(name, outspec, prms or "", binval or "", bincnt)
        return True

    def clear_modifiers(self):
        self._mods = []
        self._ctx.clear_modifiers()

    def _output_plan(self):
        """Output channels in order of first use: modifiers that name the same file share its rows (their bins side
        by side, RcontribSimulManager.cpp AddModifier / getOutput); `%s` in the name stands for the modifier."""
        plan, col0 = [], 0
        for modn, outspec, prms, binval, bincnt, ops in self._mods:
            name = (outspec or "").replace("%s", modn)
            op = next((o for o in plan if o._name == name), None)
            if op is None:
                op = RcontribOutput(name, self.get_row_max(), 0, self._dtype)
                op.omod = modn if (outspec and "%s" in outspec) else None
                plan.append(op)
            op.cols.append((col0, int(bincnt)))
            op.ncols += int(bincnt)
            op.row_bytes = op.ncols * 3 * op.dtype.itemsize
            col0 += int(bincnt)
        return plan

    def get_output(self, nm=None):
        """The named output channel, or the first one (the head of the reference's list)."""
        outs = self._outputs or self._output_plan()
        if nm is None:
            return outs[0] if outs else None
        return next((o for o in outs if o._name == nm), None)

    def get_row_max(self) -> int:
        return self.yres * (self.xres if self.xres else 1) if self.yres else 0

    def get_row_count(self) -> int:
        return self._rows_done

    def get_row_finished(self) -> int:
        return self._rows_done

    def prep_output(self) -> int:
        if not self._mods:
            raise RuntimeError("missing required modifier argument")
        try:
            self._ctx.clear_modifiers()
            for modn, outspec, prms, binval, bincnt, ops in self._mods:
                c2 = self._ctx
                # replay this modifier's cal context, then register it
                for op, arg in ops:
                    if op == "load":
                        c2.cal_load(arg)
                    else:
                        c2.cal_set(arg)
                c2.add_modifier(modn, prms, binval, bincnt)
        except RBError as e:
            raise RuntimeError(str(e)) from e
        nrows = self.get_row_max()
        for o in self._outputs:
            o.close()
        self._outputs = self._output_plan()
        done = None
        for o in self._outputs:                 # RECOVER: resume after the rows every channel already holds
            d = o.open(self, self.out_op)
            done = d if done is None else min(done, d)
        self._out = np.zeros((max(nrows, 0), self._ctx.num_columns() * 3), dtype=self._dtype)
        self._rows_done = done or 0
        return self._rows_done

    def ready(self) -> bool:
        return self._loaded and self._out is not None

    def set_thread_count(self, nt: int = 0) -> int:
        return max(1, nt)

    def n_threads(self) -> int:
        return 1

    def _flags_abi(self):
        f = 0
        if self._flags & RTimmIrrad:
            f |= _lib.RB_IRRAD_MANAGER
        if self._flags & RTlimDist:
            f |= _lib.RB_FLAG_LIMDIST
        if self._flags & RCcontrib:
            f |= _lib.RB_FLAG_CONTRIB
        return f

    def rcontrib(self, rays):
        """rays: float64 [2*nrows*accum, 3], origin and direction rows alternating."""
        if self._out is None:
            self.prep_output()
        od = np.ascontiguousarray(rays, dtype=np.float64).reshape(-1, 6)
        self._ctx.set_params(_params()._p)
        nrec = (od.shape[0] + self.accum - 1) // self.accum
        if self._out.shape[0] < nrec:
            self._out = np.zeros((nrec, self._out.shape[1]), dtype=self._dtype)
        try:
            m = self._ctx.rcontrib(od, accum=self.accum, flags=self._flags_abi(), dtype=self._dtype)
        except RBError as e:
            raise RuntimeError(str(e)) from e
        self._out[:nrec] = m.reshape(nrec, -1)
        self._rows_done = nrec
        self._store_rows(0, nrec)

    def compute_record(self, orig_direc) -> int:
        od = np.ascontiguousarray(orig_direc, dtype=np.float64).reshape(-1, 6)
        if self._out is None:
            self.prep_output()
        self._ctx.set_params(_params()._p)
        row = self._rows_done
        if row >= self._out.shape[0]:
            self._out = np.concatenate([self._out, np.zeros_like(self._out[:max(1, row)])])
        try:
            m = self._ctx.rcontrib(od[:self.accum], accum=self.accum, flags=self._flags_abi(), row_base=row,
                                   dtype=self._dtype)
        except RBError as e:
            raise RuntimeError(str(e)) from e
        self._out[row] = m.reshape(-1)
        self._rows_done += 1
        self._store_rows(row, row + 1)
        return 1

    def _store_rows(self, r0, r1):
        """Rows [r0, r1) of the context's result into the output channels (files or memory)."""
        for o in self._outputs:
            if o.array is None:
                continue
            if o.array.shape[0] < r1:
                if o._mm is not None:
                    raise RuntimeError(f"output '{o.get_name()}' holds {o.array.shape[0]} rows (set yres before prep_output)")
                o.array = np.concatenate([o.array, np.zeros((r1 - o.array.shape[0], o.array.shape[1]), dtype=o.dtype)])
            c = 0
            for col0, n in o.cols:
                o.array[r0:r1, 3 * c:3 * (c + n)] = self._out[r0:r1, 3 * col0:3 * (col0 + n)]
                c += n
            o.rows_done(self._rows_done)

    def flush_queue(self) -> int:
        return 0

    def reset_row(self, r: int) -> bool:
        self._rows_done = r
        return True

    def get_output_array(self, nm=None) -> np.ndarray:
        """[nRows, ncols * 3] view of the named channel's data (the first channel by default), as the reference's
        binding exposes the mapped file (radiance_ext.cpp:357-361)."""
        if self._out is None:
            raise RuntimeError("no output prepared")
        o = self.get_output(nm)
        if o is None or o.array is None:
            return self._out
        return o.array

    def cleanup(self, everything: bool = False) -> int:
        for o in self._outputs:
            o.close()
        self._outputs = []
        if everything:
            self._ctx.close()
            self._loaded = False
        return 0
