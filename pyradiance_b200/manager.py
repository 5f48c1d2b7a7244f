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
    def __init__(self, name, nrows, ncols):
        self._name = name
        self.n_rows = nrows
        self.row_bytes = ncols * 3 * 4
        self.cur_row = 0

    def get_name(self):
        return self._name


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
        self._rows_done = 0
        self._device = device
        if octn:
            self.load_octree(octn)

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
        return True

    def add_modifier(self, modn, outspec, prms="", binval="", bincnt=1) -> bool:
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

    def get_output(self, nm=None):
        ncols = sum(m[4] for m in self._mods)
        return RcontribOutput(nm or (self._mods[0][1] if self._mods else None), self.get_row_max(), ncols)

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
        self._out = np.zeros((max(nrows, 0), self._ctx.num_columns() * 3), dtype=self._dtype)
        self._rows_done = 0
        return nrows

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
        return 1

    def flush_queue(self) -> int:
        return 0

    def reset_row(self, r: int) -> bool:
        self._rows_done = r
        return True

    def get_output_array(self, nm=None) -> np.ndarray:
        if self._out is None:
            raise RuntimeError("no output prepared")
        return self._out

    def cleanup(self, everything: bool = False) -> int:
        if everything:
            self._ctx.close()
            self._loaded = False
        return 0
