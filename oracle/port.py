"""TEST INFRASTRUCTURE -- ctypes binding of oracle/liboracle.so (rb_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import
this module, and only as the checker."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
LIB = HERE / "liboracle.so"


class orc_params(C.Structure):
    _fields_ = [("ambounce", C.c_int), ("ambdiv", C.c_int), ("maxdepth", C.c_int), ("backvis", C.c_int),
                ("directvis", C.c_int), ("do_irrad", C.c_int), ("minweight", C.c_double), ("dstrsrc", C.c_double),
                ("specthresh", C.c_double), ("specjitter", C.c_double), ("ambval", C.c_double * 3),
                ("contrib", C.c_int), ("seed", C.c_uint64), ("srcsizerat", C.c_double)]


RESULT_DTYPE = np.dtype([("rop", "<f8", 3), ("ron", "<f8", 3), ("rot", "<f8"), ("rod", "<f8"),
                         ("robj", "<i4"), ("omod", "<i4"), ("value", "<f8", 3)])

BIN_CONST, BIN_REINHARTB, BIN_REINHART, BIN_KLEMS_FULL, BIN_HEMI, BIN_KLEMS_HALF, BIN_KLEMS_QUARTER, BIN_SHIRCHIU = range(8)

_lib = None


def build():
    subprocess.run(["make", "-C", str(HERE), "port"], check=True, capture_output=True)


def lib():
    global _lib
    if _lib is None:
        if not LIB.exists():
            build()
        L = C.CDLL(str(LIB))
        L.orc_load.restype = C.c_void_p
        L.orc_load.argtypes = [C.c_char_p, C.c_char_p, C.c_size_t]
        L.orc_free.argtypes = [C.c_void_p]
        L.orc_num_objects.argtypes = [C.c_void_p]
        L.orc_object_name.restype = C.c_char_p
        L.orc_object_name.argtypes = [C.c_void_p, C.c_int]
        L.orc_default_params.argtypes = [C.POINTER(orc_params), C.c_int]
        L.orc_set_params.argtypes = [C.c_void_p, C.POINTER(orc_params)]
        L.orc_clear_modifiers.argtypes = [C.c_void_p]
        L.orc_add_modifier.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_int, C.POINTER(C.c_double),
                                       C.POINTER(C.c_double), C.c_double, C.c_int]
        L.orc_num_columns.argtypes = [C.c_void_p]
        L.orc_rtrace.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]
        L.orc_rcontrib.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_void_p]
        L.orc_bin.restype = C.c_double
        L.orc_bin.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_double,
                              C.POINTER(C.c_double)]
        L.orc_get_counters.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_reset_counters.argtypes = [C.c_void_p]
        L.orc_set_walker.argtypes = [C.c_void_p, C.c_int]
        L.orc_dev_nodes.restype = C.c_uint64
        L.orc_dev_nodes.argtypes = [C.c_void_p]
        L.orc_last_error.restype = C.c_char_p
        L.orc_last_error.argtypes = [C.c_void_p]
        _lib = L
    return _lib


def _v3(v):
    return (C.c_double * 3)(*[float(x) for x in v])


def bin_of(fn, mf, n, u, rhs, d):
    return lib().orc_bin(fn, mf, _v3(n), _v3(u), float(rhs), _v3(d))


class Scene:
    def __init__(self, octree, rcontrib=False, **params):
        L = lib()
        err = C.create_string_buffer(512)
        self.h = L.orc_load(os.fspath(octree).encode(), err, 512)
        if not self.h:
            raise RuntimeError("oracle: " + err.value.decode())
        self.p = orc_params()
        L.orc_default_params(C.byref(self.p), 1 if rcontrib else 0)
        self.set(**params)

    def set(self, **kw):
        for k, v in kw.items():
            if k == "ambval":
                for i in range(3):
                    self.p.ambval[i] = v[i]
            else:
                setattr(self.p, k, v)
        lib().orc_set_params(self.h, C.byref(self.p))

    def set_walker(self, mode):
        """1: the device's integer walk restated on the CPU; 0: the recursive restatement of the reference."""
        lib().orc_set_walker(self.h, int(mode))

    def close(self):
        if self.h:
            lib().orc_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def name(self, i):
        return lib().orc_object_name(self.h, int(i)).decode() if i >= 0 else "*"

    def add_modifier(self, name, fn=BIN_CONST, mf=1, n=(0, 0, -1), u=(0, 1, 0), rhs=1.0, nbins=1):
        return lib().orc_add_modifier(self.h, name.encode(), fn, mf, _v3(n), _v3(u), rhs, nbins)

    def clear_modifiers(self):
        lib().orc_clear_modifiers(self.h)

    def ncols(self):
        return lib().orc_num_columns(self.h)

    def rtrace(self, rays, irrad=0):
        rays = np.ascontiguousarray(rays, dtype=np.float64).reshape(-1, 6)
        out = np.zeros(rays.shape[0], dtype=RESULT_DTYPE)
        if lib().orc_rtrace(self.h, rays.ctypes.data, rays.shape[0], irrad, out.ctypes.data) < 0:
            raise RuntimeError("oracle: " + lib().orc_last_error(self.h).decode())
        return out

    def rcontrib(self, rays, accum=1, irrad=0):
        rays = np.ascontiguousarray(rays, dtype=np.float64).reshape(-1, 6)
        nrec = (rays.shape[0] + accum - 1) // accum
        out = np.zeros((nrec, self.ncols(), 3), dtype=np.float64)
        if lib().orc_rcontrib(self.h, rays.ctypes.data, rays.shape[0], accum, irrad, out.ctypes.data) < 0:
            raise RuntimeError("oracle: " + lib().orc_last_error(self.h).decode())
        return out

    def counters(self):
        c = (C.c_uint64 * 5)()
        lib().orc_get_counters(self.h, c)
        return dict(zip(["nrays", "nodes", "leafents", "prims", "contribs"], [int(x) for x in c]))

    def reset_counters(self):
        lib().orc_reset_counters(self.h)
