"""ctypes binding of libqtt_b200.so (include/qtt_b200.h).  There is NO fallback: if the shared
library is missing or no sm_100 device is usable, every entry point raises."""
import ctypes as C
import os

import numpy as np

__all__ = ["lib", "lib_path", "QttError", "SweepParams", "SweepInfo", "Context", "default_context", "check",
           "EXPORTED_SYMBOLS"]

_HERE = os.path.dirname(os.path.abspath(__file__))
lib_path = os.path.join(_HERE, "libqtt_b200.so")


class QttError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libqtt_b200 error {code}: {msg}")
        self.code = code


class SweepParams(C.Structure):
    _fields_ = [
        ("nsweeps", C.c_int32), ("maxdim", C.POINTER(C.c_int32)), ("mindim", C.POINTER(C.c_int32)),
        ("cutoff", C.POINTER(C.c_double)), ("solver", C.c_int32), ("a0", C.c_double), ("a1", C.c_double),
        ("tol", C.c_double), ("krylovdim", C.c_int32), ("maxiter", C.c_int32), ("fixed_matvecs", C.c_int32),
        ("normalize", C.c_int32), ("outputlevel", C.c_int32), ("target", C.c_double),
        ("svd_max_sweeps", C.c_int32), ("x0_canonical", C.c_int32),
    ]


class SweepInfo(C.Structure):
    _fields_ = [
        ("maxlinkdim", C.c_int32), ("svd_sweeps_max", C.c_int32), ("maxtruncerr", C.c_double), ("seconds", C.c_double),
        ("solver_resid", C.c_double), ("matvecs", C.c_int64), ("flops_matvec", C.c_double), ("t_matvec", C.c_double),
        ("t_env", C.c_double), ("t_svd", C.c_double), ("t_krylov", C.c_double), ("energy", C.c_double),
        ("device_ms", C.c_double),
    ]

    def as_dict(self):
        return {f: getattr(self, f) for f, _ in self._fields_}


_vp = C.c_void_p
_dp = C.POINTER(C.c_double)
_i64p = C.POINTER(C.c_int64)
_vpp = C.POINTER(C.c_void_p)

# name -> (restype, argtypes); the `-m "not gpu"` tests check that every one of these is exported
EXPORTED_SYMBOLS = {
    "qtt_version": (C.c_int, []),
    "qtt_ctx_create": (_vp, [C.c_int]),
    "qtt_ctx_destroy": (None, [_vp]),
    "qtt_last_error": (C.c_char_p, [_vp]),
    "qtt_global_error": (C.c_char_p, []),
    "qtt_ctx_sync": (C.c_int, [_vp]),
    "qtt_ctx_launch_count": (C.c_int64, [_vp]),
    "qtt_ctx_device_info": (C.c_int, [_vp, _i64p]),
    "qtt_mps_upload": (C.c_int, [_vp, C.c_int, _i64p, _vpp, C.c_int, _vpp]),
    "qtt_mpo_upload": (C.c_int, [_vp, C.c_int, _i64p, _vpp, C.c_int, _vpp]),
    "qtt_mps_length": (C.c_int, [_vp]),
    "qtt_mps_dtype": (C.c_int, [_vp]),
    "qtt_mps_dims": (C.c_int, [_vp, _i64p]),
    "qtt_mps_download": (C.c_int, [_vp, _vpp]),
    "qtt_mps_clone": (C.c_int, [_vp, _vpp]),
    "qtt_mpo_length": (C.c_int, [_vp]),
    "qtt_mpo_dims": (C.c_int, [_vp, _i64p]),
    "qtt_mps_free": (None, [_vp]),
    "qtt_mpo_free": (None, [_vp]),
    "qtt_linsolve": (C.c_int, [_vp, _vp, _vp, _vp, C.POINTER(SweepParams), C.POINTER(SweepInfo)]),
    "qtt_dmrg_target": (C.c_int, [_vp, _vp, _vp, C.POINTER(SweepParams), C.POINTER(SweepInfo)]),
    "qtt_set_profiling": (C.c_int, [_vp, C.c_int]),
    "qtt_apply": (C.c_int, [_vp, _vp, _vp, C.c_double, C.c_int32, C.c_int32, _vpp]),
    "qtt_vec_to_mps": (C.c_int, [_vp, C.c_void_p, C.c_int32, C.c_int32, C.c_double, C.c_int32, C.c_int32, _vpp, _dp]),
    "qtt_mps_add": (C.c_int, [_vp, _vp, _vp, C.c_double, C.c_double, C.c_double, C.c_int32, C.c_int32, _vpp]),
    "qtt_mps_to_vec": (C.c_int, [_vp, _vp, C.c_void_p, C.c_int32]),
    "qtt_mps_sample": (C.c_int, [_vp, _vp, _i64p, C.c_int64, C.c_void_p]),
    "qtt_inner": (C.c_int, [_vp, _vp, _vp, _dp]),
    "qtt_residual": (C.c_int, [_vp, _vp, _vp, _vp, _dp, _dp]),
    "qtt_matvec2": (C.c_int, [_vp, _dp, _dp, _dp, _dp, _dp, _dp, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int64,
                              C.c_int32, C.POINTER(C.c_float)]),
    "qtt_env_left": (C.c_int, [_vp, _dp, _dp, _dp, _dp, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int32,
                               C.POINTER(C.c_float)]),
    "qtt_env_right": (C.c_int, [_vp, _dp, _dp, _dp, _dp, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int32,
                                C.POINTER(C.c_float)]),
    "qtt_split2": (C.c_int, [_vp, _dp, C.c_int64, C.c_int64, C.c_int32, C.c_double, C.c_int32, C.c_int32, _dp, _dp, _dp,
                             C.POINTER(C.c_int32), _dp, C.c_int32, C.POINTER(C.c_float)]),
    "qtt_gemm": (C.c_int, [_vp, C.c_int32, C.c_int32, C.c_int64, C.c_int64, C.c_int64, C.c_double, _dp, C.c_int64, _dp,
                           C.c_int64, C.c_double, _dp, C.c_int64, C.c_int32, C.c_int32, C.POINTER(C.c_float)]),
    "qtt_krylov_op": (C.c_int, [_vp, C.c_int32, _dp, C.c_int64, C.c_int64, _dp, _dp, _dp, C.c_int32, C.POINTER(C.c_float)]),
}


def _load():
    if not os.path.exists(lib_path):
        raise ImportError(
            f"{lib_path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(libqtt_b200 has no CPU fallback)")
    L = C.CDLL(lib_path)
    for name, (res, args) in EXPORTED_SYMBOLS.items():
        fn = getattr(L, name)  # AttributeError if a declared symbol is not exported
        fn.restype = res
        fn.argtypes = args
    return L


lib = _load()


def check(ctx_handle, rc):
    if rc != 0:
        msg = lib.qtt_last_error(ctx_handle) if ctx_handle else lib.qtt_global_error()
        raise QttError(rc, (msg or b"").decode("utf-8", "replace"))


class Context:
    """One CUDA device + stream + workspaces (`qtt_ctx`).  Not thread-safe: one per host thread / GPU."""

    def __init__(self, device=0):
        self.handle = lib.qtt_ctx_create(int(device))
        if not self.handle:
            raise QttError(-2, (lib.qtt_global_error() or b"").decode("utf-8", "replace"))
        self.device = device

    def close(self):
        if getattr(self, "handle", None):
            lib.qtt_ctx_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sync(self):
        check(self.handle, lib.qtt_ctx_sync(self.handle))

    @property
    def launches(self):
        return int(lib.qtt_ctx_launch_count(self.handle))

    def device_info(self):
        out = (C.c_int64 * 4)()
        check(self.handle, lib.qtt_ctx_device_info(self.handle, out))
        return dict(sms=out[0], major=out[1], minor=out[2], clock_khz=out[3])

    def set_profiling(self, on):
        check(self.handle, lib.qtt_set_profiling(self.handle, int(bool(on))))


_default = {}


def default_context(device=None):
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", "0"))
    if device not in _default:
        _default[device] = Context(device)
    return _default[device]


def as_f(a, dtype=np.float64):
    """Column-major (Julia layout) contiguous copy."""
    return np.asfortranarray(np.asarray(a, dtype=dtype))


def dptr(a):
    return a.ctypes.data_as(_dp)
