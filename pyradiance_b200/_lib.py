"""ctypes binding of librb200.so (the C ABI declared in include/rb200.h).

This is the stand-in for the reference's nanobind module
(src/binding/radiance_ext.cpp): numpy in, numpy out, no PyTorch.  The shared
library is built in-tree by ``__graft_entry__.build()`` / ``csrc/Makefile``.
There is no CPU fallback: if the library is missing, or no CUDA device is
present, every compute call raises.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
LIB_PATH = _HERE / "librb200.so"


class rb_params(C.Structure):
    _fields_ = [
        ("do_irrad", C.c_int), ("rand_samp", C.c_int), ("dstrsrc", C.c_double),
        ("shadthresh", C.c_double), ("shadcert", C.c_double), ("directrelay", C.c_int),
        ("vspretest", C.c_int), ("directvis", C.c_int), ("srcsizerat", C.c_double),
        ("cextinction", C.c_double * 3), ("salbedo", C.c_double * 3), ("seccg", C.c_double),
        ("ssampdist", C.c_double), ("specthresh", C.c_double), ("specjitter", C.c_double),
        ("backvis", C.c_int), ("maxdepth", C.c_int), ("minweight", C.c_double),
        ("ambval", C.c_double * 3), ("ambvwt", C.c_int), ("ambacc", C.c_double),
        ("ambres", C.c_int), ("ambdiv", C.c_int), ("ambssamp", C.c_int), ("ambounce", C.c_int),
    ]


class rb_view(C.Structure):
    _fields_ = [("type", C.c_int), ("vp", C.c_double * 3), ("vdir", C.c_double * 3), ("hvec", C.c_double * 3),
                ("vvec", C.c_double * 3), ("horiz", C.c_double), ("vert", C.c_double), ("hoff", C.c_double),
                ("voff", C.c_double), ("vfore", C.c_double), ("vaft", C.c_double), ("hn2", C.c_double), ("vn2", C.c_double)]


class rb_stats(C.Structure):
    _fields_ = [
        ("nrays", C.c_uint64), ("nodes", C.c_uint64), ("leafents", C.c_uint64),
        ("prims", C.c_uint64), ("contribs", C.c_uint64), ("launches", C.c_uint64),
        ("wave_launches", C.c_uint64), ("waves", C.c_uint64), ("batches", C.c_uint64),
        ("retries", C.c_uint64), ("badbin", C.c_uint64), ("kernel_ms", C.c_double),
        ("wave_ms", C.c_double), ("shade_ms", C.c_double),
    ]


RAY_RESULT_DTYPE = np.dtype([
    ("rop", "<f8", 3), ("ron", "<f8", 3), ("rot", "<f8"), ("rod", "<f8"),
    ("robj", "<i4"), ("omod", "<i4"), ("rweight", "<f4"), ("pad", "<i4"), ("pert", "<f8", (3,)),
])

RB_IRRAD_NONE, RB_IRRAD_RTRACE, RB_IRRAD_RCONTRIB, RB_IRRAD_MANAGER = 0, 1, 2, 3
RB_FLAG_LIMDIST = 4
RB_FLAG_CONTRIB = 8
RB_FLAG_RAYS_ON_DEVICE = 16
RB_FLAG_OUT_ON_DEVICE = 32
RB_FLAG_OUT_DOUBLE = 64
RB_PROGRAM_RTRACE, RB_PROGRAM_RCONTRIB = 0, 1

# every symbol include/rb200.h declares: (name, restype, argtypes)
_P = C.c_void_p
SYMBOLS = [
    ("rb_create", _P, [C.c_int]),
    ("rb_destroy", None, [_P]),
    ("rb_last_error", C.c_char_p, [_P]),
    ("rb_version", C.c_char_p, []),
    ("rb_device_count", C.c_int, []),
    ("rb_set_row_base", C.c_int, [C.c_void_p, C.c_uint64]),
    ("rb_set_defaults", C.c_int, [_P, C.c_int]),
    ("rb_get_params", C.c_int, [_P, C.POINTER(rb_params)]),
    ("rb_set_params", C.c_int, [_P, C.POINTER(rb_params)]),
    ("rb_set_option", C.c_int, [_P, C.c_int, C.POINTER(C.c_char_p)]),
    ("rb_load_octree", C.c_int, [_P, C.c_char_p]),
    ("rb_save_octree", C.c_int, [_P, C.c_char_p]),
    ("rb_num_objects", C.c_int, [_P]),
    ("rb_object_name", C.c_char_p, [_P, C.c_int]),
    ("rb_object_type", C.c_char_p, [_P, C.c_int]),
    ("rb_object_modifier", C.c_int, [_P, C.c_int]),
    ("rb_num_header_lines", C.c_int, [_P]),
    ("rb_header_line", C.c_char_p, [_P, C.c_int]),
    ("rb_scene_warnings", C.c_char_p, [_P]),
    ("rb_cal_load", C.c_int, [_P, C.c_char_p]),
    ("rb_cal_set", C.c_int, [_P, C.c_char_p]),
    ("rb_cal_eval", C.c_int, [_P, C.c_char_p, C.POINTER(C.c_double)]),
    ("rb_clear_modifiers", C.c_int, [_P]),
    ("rb_add_modifier", C.c_int, [_P, C.c_char_p, C.c_char_p, C.c_char_p, C.c_int]),
    ("rb_num_columns", C.c_int, [_P]),
    ("rb_bin_of_direction", C.c_int, [_P, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    ("rb_rcontrib", C.c_int, [_P, _P, C.c_size_t, C.c_int, C.c_uint, C.c_uint64, _P, C.c_size_t]),
    ("rb_rtrace", C.c_int, [_P, _P, C.c_size_t, C.c_uint, _P, _P]),
    ("rb_get_stats", C.c_int, [_P, C.POINTER(rb_stats)]),
    ("rb_reset_stats", C.c_int, [_P]),
    ("rb_set_stream", C.c_int, [_P, _P]),
    ("rb_set_seed", C.c_int, [_P, C.c_uint64]),
    ("rb_set_queue_capacity", C.c_int, [_P, C.c_size_t]),
    ("rb_device_alloc", _P, [_P, C.c_size_t]),
    ("rb_device_free", C.c_int, [_P, _P]),
    ("rb_device_upload", C.c_int, [_P, _P, _P, C.c_size_t]),
    ("rb_device_download", C.c_int, [_P, _P, _P, C.c_size_t]),
    ("rb_device_sync", C.c_int, [_P]),
    ("rb_host_register", C.c_int, [_P, _P, C.c_size_t]),
    ("rb_host_unregister", C.c_int, [_P, _P]),
    ("rb_view_rays", C.c_int, [_P, C.POINTER(rb_view), C.c_int, C.c_int, C.c_int, C.c_double, C.c_uint64, _P, C.c_uint]),
    ("rb_ipc_export", C.c_int, [_P, _P, _P]),
    ("rb_ipc_open", C.c_int, [_P, _P, C.POINTER(C.c_void_p)]),
    ("rb_ipc_close", C.c_int, [_P, _P]),
    ("rb_oconv", C.c_int, [C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_char_p, C.c_size_t]),
    ("rb_mtx_multiply", C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_uint,
                                  C.POINTER(C.c_double)]),
    ("rb_format_ascii", C.c_size_t, [C.c_void_p, C.c_int, C.c_size_t, C.c_size_t, C.c_int, C.c_void_p, C.c_size_t]),
    ("rb_oconv_files", C.c_int, [C.POINTER(C.c_char_p), C.c_int, C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_char_p,
                                 C.c_size_t]),
]

_lib = None


def load_library() -> C.CDLL:
    """Load librb200.so; raises (loudly) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("RB200_LIBRARY", str(LIB_PATH))
    if not os.path.exists(path):
        raise RuntimeError(
            f"{path} not found: build the CUDA library first "
            "(python -c 'import __graft_entry__ as g; g.build()' or make -C pyradiance_b200/csrc). "
            "pyradiance_b200 has no CPU fallback.")
    lib = C.CDLL(path)
    for name, res, args in SYMBOLS:
        fn = getattr(lib, name)       # AttributeError if the ABI is incomplete
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class RBError(RuntimeError):
    """Failure reported by the C ABI (the reference exits the process instead:
    common/error.c:24-46; pyradiance turns that into RuntimeError, anci.py:13-30)."""


class Context:
    """Thin object wrapper of rb_ctx."""

    def __init__(self, device: int = 0, program: int = RB_PROGRAM_RTRACE):
        self.lib = load_library()
        self.h = self.lib.rb_create(int(device))
        if not self.h:
            raise RBError("rb_create failed")
        self.lib.rb_set_defaults(self.h, program)
        self.device = device

    def close(self):
        if getattr(self, "h", None):
            self.lib.rb_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rv):
        if rv < 0:
            raise RBError(self.lib.rb_last_error(self.h).decode("utf-8", "replace"))
        return rv

    # ---- options ----
    def set_defaults(self, program):
        self._ck(self.lib.rb_set_defaults(self.h, program))

    def get_params(self) -> rb_params:
        p = rb_params()
        self._ck(self.lib.rb_get_params(self.h, C.byref(p)))
        return p

    def set_params(self, p: rb_params):
        self._ck(self.lib.rb_set_params(self.h, C.byref(p)))

    def set_option(self, argv) -> int:
        """Parse one option at argv[0]; returns extra args consumed or -1."""
        arr = (C.c_char_p * len(argv))(*[a.encode() for a in argv])
        rv = self.lib.rb_set_option(self.h, len(argv), arr)
        if rv == -2:
            self._ck(-1)
        return rv

    def set_options(self, argv):
        """Consume a whole list of render options; raises on an unknown one."""
        i = 0
        argv = list(argv)
        while i < len(argv):
            rv = self.set_option(argv[i:])
            if rv < 0:
                raise RBError(f"bad or unsupported option '{argv[i]}'")
            i += rv + 1

    # ---- scene ----
    def load_octree(self, path):
        self._ck(self.lib.rb_load_octree(self.h, os.fspath(path).encode()))

    def parse_octree(self, path):
        """Parse + flatten only (works without a GPU); returns the error the
        GPU upload would have given, or None."""
        rv = self.lib.rb_load_octree(self.h, os.fspath(path).encode())
        if rv < 0:
            return self.lib.rb_last_error(self.h).decode()
        return None

    def save_octree(self, path):
        """Write the loaded scene (instances expanded) as a frozen octree."""
        self._ck(self.lib.rb_save_octree(self.h, os.fspath(path).encode()))

    def num_objects(self):
        return self.lib.rb_num_objects(self.h)

    def object_name(self, i):
        return self.lib.rb_object_name(self.h, int(i)).decode()

    def object_type(self, i):
        return self.lib.rb_object_type(self.h, int(i)).decode()

    def object_modifier(self, i):
        return self.lib.rb_object_modifier(self.h, int(i))

    def header_lines(self):
        return [self.lib.rb_header_line(self.h, i).decode("latin-1")
                for i in range(self.lib.rb_num_header_lines(self.h))]

    def warnings(self):
        return self.lib.rb_scene_warnings(self.h).decode()

    # ---- cal context / modifiers ----
    def cal_load(self, fname):
        self._ck(self.lib.rb_cal_load(self.h, fname.encode()))

    def cal_set(self, assignments):
        self._ck(self.lib.rb_cal_set(self.h, assignments.encode()))

    def cal_eval(self, expr) -> float:
        v = C.c_double()
        self._ck(self.lib.rb_cal_eval(self.h, expr.encode(), C.byref(v)))
        return v.value

    def clear_modifiers(self):
        self._ck(self.lib.rb_clear_modifiers(self.h))

    def add_modifier(self, modname, params="", binexpr="0", nbins=1) -> int:
        return self._ck(self.lib.rb_add_modifier(
            self.h, modname.encode(), (params or "").encode(), (binexpr or "0").encode(), int(nbins)))

    def num_columns(self):
        return self.lib.rb_num_columns(self.h)

    def bin_of_direction(self, modifier_index, d) -> float:
        v = C.c_double()
        dd = (C.c_double * 3)(*[float(x) for x in d])
        self._ck(self.lib.rb_bin_of_direction(self.h, int(modifier_index), dd, C.byref(v)))
        return v.value

    # ---- compute ----
    def rcontrib(self, rays, accum=1, flags=RB_IRRAD_NONE, row_base=0, out=None, dtype=np.float32):
        rays = np.ascontiguousarray(rays, dtype=np.float64).reshape(-1, 6)
        n = rays.shape[0]
        nrec = (n + accum - 1) // accum
        ncols = self.num_columns()
        if out is None:
            out = np.empty((nrec, ncols, 3), dtype=dtype)
        assert out.dtype in (np.float32, np.float64) and out.flags["C_CONTIGUOUS"] and out.size >= nrec * ncols * 3
        if out.dtype == np.float64:
            flags |= RB_FLAG_OUT_DOUBLE
        self._ck(self.lib.rb_rcontrib(self.h, rays.ctypes.data, n, int(accum), int(flags), int(row_base),
                                      out.ctypes.data, out.size))
        return out

    def mtx_multiply(self, a: np.ndarray, b: np.ndarray) -> np.ndarray:
        """[nr, ni, 3] x [ni, nc, 3] -> [nr, nc, 3] float32, per colour channel (C ABI rb_mtx_multiply)."""
        a = np.ascontiguousarray(a, dtype=np.float32)
        b = np.ascontiguousarray(b, dtype=np.float32)
        assert a.ndim == 3 and b.ndim == 3 and a.shape[2] == 3 and b.shape[2] == 3 and a.shape[1] == b.shape[0]
        out = np.empty((a.shape[0], b.shape[1], 3), dtype=np.float32)
        ms = C.c_double(0)
        self._ck(self.lib.rb_mtx_multiply(self.h, a.ctypes.data, a.shape[0], a.shape[1], b.ctypes.data, b.shape[1],
                                          out.ctypes.data, 0, C.byref(ms)))
        self.last_mtx_ms = ms.value
        return out

    def rcontrib_device(self, d_rays, nrays, accum, flags, row_base, d_out, out_floats):
        """Both buffers already in HBM (raw device pointers)."""
        self._ck(self.lib.rb_rcontrib(self.h, d_rays, int(nrays), int(accum),
                                      int(flags) | RB_FLAG_RAYS_ON_DEVICE | RB_FLAG_OUT_ON_DEVICE,
                                      int(row_base), d_out, int(out_floats)))

    def rtrace(self, rays, flags=RB_IRRAD_NONE, want_values=True, want_results=True, row_base=0):
        rays = np.ascontiguousarray(rays, dtype=np.float64).reshape(-1, 6)
        self._ck(self.lib.rb_set_row_base(self.h, int(row_base)))
        n = rays.shape[0]
        values = np.zeros((n, 3), dtype=np.float64) if want_values else None
        results = np.zeros(n, dtype=RAY_RESULT_DTYPE) if want_results else None
        self._ck(self.lib.rb_rtrace(self.h, rays.ctypes.data, n, int(flags),
                                    values.ctypes.data if want_values else None,
                                    results.ctypes.data if want_results else None))
        return values, results

    def stats(self) -> dict:
        s = rb_stats()
        self.lib.rb_get_stats(self.h, C.byref(s))
        return {k: getattr(s, k) for k, _ in rb_stats._fields_}

    def reset_stats(self):
        self.lib.rb_reset_stats(self.h)

    def set_stream(self, stream_ptr):
        self._ck(self.lib.rb_set_stream(self.h, stream_ptr))

    def set_seed(self, seed):
        self.lib.rb_set_seed(self.h, int(seed) & 0xFFFFFFFFFFFFFFFF)

    def set_queue_capacity(self, nrays):
        self.lib.rb_set_queue_capacity(self.h, int(nrays))

    def device_alloc(self, nbytes):
        p = self.lib.rb_device_alloc(self.h, int(nbytes))
        if not p:
            self._ck(-1)
        return p

    def device_free(self, p):
        self.lib.rb_device_free(self.h, p)

    def device_upload(self, dptr, arr):
        arr = np.ascontiguousarray(arr)
        self._ck(self.lib.rb_device_upload(self.h, dptr, arr.ctypes.data, arr.nbytes))

    def device_download(self, arr, dptr):
        self._ck(self.lib.rb_device_download(self.h, arr.ctypes.data, dptr, arr.nbytes))

    def pin(self, arr):
        """Page-lock a numpy array in place (cudaHostRegister)."""
        self._ck(self.lib.rb_host_register(self.h, arr.ctypes.data, arr.nbytes))

    def unpin(self, arr):
        self.lib.rb_host_unregister(self.h, arr.ctypes.data)

    def sync(self):
        self._ck(self.lib.rb_device_sync(self.h))

    # ---- view rays on the device (C ABI rb_view_rays) ----
    def view_rays(self, view, xres, yres, repeat=1, pj=0.0, seed=0, out_ptr=None):
        """`view`: a pyradiance_b200.views.View after set_view().  Returns a float64 [n, 6] array, or -- with
        out_ptr, a device pointer with room for xres * yres * repeat rays -- writes there and returns the count."""
        v = rb_view()
        v.type = ord(view.type)
        for k in range(3):
            v.vp[k], v.vdir[k], v.hvec[k], v.vvec[k] = float(view.vp[k]), float(view.vdir[k]), float(view.hvec[k]), float(view.vvec[k])
        v.horiz, v.vert, v.hoff, v.voff = view.horiz, view.vert, view.hoff, view.voff
        v.vfore, v.vaft, v.hn2, v.vn2 = view.vfore, view.vaft, view.hn2, view.vn2
        n = int(xres) * int(yres) * int(repeat)
        if out_ptr is not None:
            self._ck(self.lib.rb_view_rays(self.h, C.byref(v), int(xres), int(yres), int(repeat), float(pj), int(seed), out_ptr,
                                           RB_FLAG_OUT_ON_DEVICE))
            return n
        out = np.empty((n, 6), dtype=np.float64)
        self._ck(self.lib.rb_view_rays(self.h, C.byref(v), int(xres), int(yres), int(repeat), float(pj), int(seed),
                                       out.ctypes.data, 0))
        return out

    # ---- peer-memory window (the row gather of SURVEY 8e, see include/rb200.h) ----
    def ipc_export(self, dptr) -> bytes:
        buf = C.create_string_buffer(64)
        self._ck(self.lib.rb_ipc_export(self.h, dptr, buf))
        return buf.raw

    def ipc_open(self, handle: bytes):
        p = C.c_void_p(0)
        self._ck(self.lib.rb_ipc_open(self.h, C.create_string_buffer(handle, 64), C.byref(p)))
        return p.value

    def ipc_close(self, dptr):
        self._ck(self.lib.rb_ipc_close(self.h, dptr))


def device_count() -> int:
    """CUDA devices visible to this process (C ABI rb_device_count)."""
    return int(load_library().rb_device_count())


def format_ascii(values: np.ndarray, triplets: bool = False) -> bytes:
    """[nrows, ...] float32/float64 -> the "%e\\t" ... "\\n" text rtrace / rcontrib write (C ABI rb_format_ascii)."""
    v = np.ascontiguousarray(values)
    if v.dtype not in (np.float32, np.float64):
        v = v.astype(np.float64)
    nrows = v.shape[0] if v.ndim > 0 else 0
    if nrows == 0:
        return b""
    per_row = v.size // nrows
    lib = load_library()
    buf = bytearray(v.size * 16 + nrows)
    cbuf = (C.c_char * len(buf)).from_buffer(buf)
    n = lib.rb_format_ascii(v.ctypes.data, int(v.dtype == np.float64), nrows, per_row, int(triplets), C.addressof(cbuf),
                            len(buf))
    assert n <= len(buf)
    del cbuf
    return bytes(buf[:n])


def oconv_files(rad_paths, oct_path, include_octree=None, objlim=6, maxres=16384):
    """Own octree builder, `oconv -f [-i include_octree] rad_paths...` (C ABI rb_oconv_files)."""
    lib = load_library()
    buf = C.create_string_buffer(1024)
    arr = (C.c_char_p * len(rad_paths))(*[os.fspath(p).encode() for p in rad_paths])
    rv = lib.rb_oconv_files(arr, len(rad_paths), os.fspath(include_octree).encode() if include_octree else None,
                            os.fspath(oct_path).encode(), objlim, maxres, buf, 1024)
    if rv < 0:
        raise RBError(buf.value.decode())


def oconv_file(rad_path, oct_path, objlim=6, maxres=16384):
    """Own octree builder: text scene -> frozen .oct (C ABI rb_oconv)."""
    lib = load_library()
    buf = C.create_string_buffer(1024)
    rv = lib.rb_oconv(os.fspath(rad_path).encode(), os.fspath(oct_path).encode(), objlim, maxres, buf, 1024)
    if rv < 0:
        raise RBError(buf.value.decode())
