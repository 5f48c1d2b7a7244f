"""pyradiance_b200 -- the rtrace / rcontrib hot path of LBNL-ETA/pyradiance as
hand-written CUDA for B200 (sm_100a), behind pyradiance's own API names.

Process-boundary look-alikes:  rtrace(), Rcontrib            (rt.py)
In-process look-alikes:        RtraceSimulManager, RcontribSimulManager,
                               RayParams, get_ray_params, set_ray_params,
                               set_option, initfunc, loadfunc, eval,
                               set_eparams, calcontext        (manager.py)
C ABI:                         include/rb200.h, bound in _lib.py

Importing the package does not need a GPU; every compute call does, and
raises if the CUDA library or device is missing (there is no CPU fallback).
"""
from ._lib import Context, RBError, load_library, oconv_file  # noqa: F401
from .manager import (RCCONTEXT, RCcontrib, RTdoFIFO, RTimmIrrad, RTlimDist, RTmask,  # noqa: F401
                      RTtraceSources, RayParams, RcOutputOp, RcontribSimulManager, RtraceSimulManager,
                      calcontext, eval, get_ray_params, initfunc, loadfunc, ray_done, set_eparams, set_option,
                      set_ray_params, setspectrsamp)
from .rt import Rcontrib, rcontrib_main, rtrace, rtrace_main  # noqa: F401
from .fluxmtx import rfluxmtx, rfluxmtx_main  # noqa: F401
from .views import vwrays, vwrays_main  # noqa: F401
from .mtx import Rmtxop, dctimestep, dctimestep_main, rmtxop, rmtxop_main  # noqa: F401

__version__ = "0.1.0"
