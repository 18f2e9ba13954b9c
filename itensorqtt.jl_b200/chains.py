"""Device-resident chains.  Host side uses numpy cores in the reference's logical index order:
MPS core (chi_l, 2, chi_r); MPO core (w_l, w_r, s_in, s_out) [reference src/laplacian.jl:32].
Only shapes and dense data cross the C ABI (column-major, as Julia's `array(x[j], inds...)` would)."""
import ctypes as C

import numpy as np

from ._lib import lib, check, default_context

__all__ = ["DeviceMPS", "DeviceMPO", "linkdims", "maxlinkdim"]


def _dtype_code(cores):
    return 1 if any(np.iscomplexobj(c) for c in cores) else 0


def _upload(ctx, cores, dims_of, fn):
    n = len(cores)
    code = _dtype_code(cores)
    npdt = np.complex128 if code else np.float64
    keep = [np.asfortranarray(np.asarray(c, dtype=npdt)) for c in cores]
    dims = (C.c_int64 * (n + 1))(*dims_of(cores))
    ptrs = (C.c_void_p * n)(*[k.ctypes.data for k in keep])
    out = C.c_void_p()
    check(ctx.handle, fn(ctx.handle, n, dims, ptrs, code, C.byref(out)))
    return out, code


class DeviceMPS:
    """`qtt_mps`: an MPS/QTT resident in HBM."""

    def __init__(self, cores=None, ctx=None, _handle=None):
        self.ctx = ctx or default_context()
        if _handle is not None:
            self.handle = _handle
            return
        cores = [np.asarray(c) for c in cores]
        for j, c in enumerate(cores):
            if c.ndim != 3 or c.shape[1] != 2:
                raise ValueError(f"MPS core {j} must have shape (chi_l, 2, chi_r), got {c.shape}")
        for j in range(len(cores) - 1):
            if cores[j].shape[2] != cores[j + 1].shape[0]:
                raise ValueError(f"bond mismatch between cores {j} and {j + 1}")
        self.handle, _ = _upload(self.ctx, cores, lambda cs: [cs[0].shape[0]] + [c.shape[2] for c in cs], lib.qtt_mps_upload)

    def __len__(self):
        return int(lib.qtt_mps_length(self.handle))

    @property
    def is_complex(self):
        return int(lib.qtt_mps_dtype(self.handle)) == 1

    def dims(self):
        n = len(self)
        out = (C.c_int64 * (n + 1))()
        check(self.ctx.handle, lib.qtt_mps_dims(self.handle, out))
        return list(out)

    def download(self, order="C", out=None):
        """Host cores, shape (chi_l, 2, chi_r).  order="F" returns the column-major buffers the C ABI filled (what a Julia
        caller gets: no host-side transpose); order="C" makes C-contiguous copies.  `out`: caller-owned column-major buffers
        of exactly these shapes to fill instead of fresh ones -- e.g. PINNED memory, which the device-to-host copies then
        reach at full PCIe/C2C rate (a pageable destination is staged by the driver: ~3x slower for a 16 MB chain)."""
        chi = self.dims()
        n = len(chi) - 1
        dt = np.complex128 if self.is_complex else np.float64
        if out is not None:
            if len(out) != n:
                raise ValueError(f"download(out=...): {len(out)} buffers for {n} cores")
            for j, b in enumerate(out):
                if b.shape != (chi[j], 2, chi[j + 1]) or b.dtype != dt or not b.flags.f_contiguous:
                    raise ValueError(f"download(out=...): buffer {j} must be a column-major {dt.__name__} array of shape "
                                     f"{(chi[j], 2, chi[j + 1])}, got {b.dtype} {b.shape}")
            bufs = list(out)
        else:
            bufs = [np.empty((chi[j], 2, chi[j + 1]), dtype=dt, order="F") for j in range(n)]
        ptrs = (C.c_void_p * n)(*[b.ctypes.data for b in bufs])
        check(self.ctx.handle, lib.qtt_mps_download(self.handle, ptrs))
        return bufs if order == "F" else [np.ascontiguousarray(b) for b in bufs]

    def clone(self):
        out = C.c_void_p()
        check(self.ctx.handle, lib.qtt_mps_clone(self.handle, C.byref(out)))
        return DeviceMPS(ctx=self.ctx, _handle=out)

    def free(self):
        if getattr(self, "handle", None):
            lib.qtt_mps_free(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class DeviceMPO:
    """`qtt_mpo`: an MPO resident in HBM."""

    def __init__(self, cores, ctx=None):
        self.ctx = ctx or default_context()
        cores = [np.asarray(c) for c in cores]
        for j, c in enumerate(cores):
            if c.ndim != 4 or c.shape[2:] != (2, 2):
                raise ValueError(f"MPO core {j} must have shape (w_l, w_r, 2, 2), got {c.shape}")
        self.handle, _ = _upload(self.ctx, cores, lambda cs: [cs[0].shape[0]] + [c.shape[1] for c in cs], lib.qtt_mpo_upload)
        self.n = len(cores)

    def __len__(self):
        return self.n

    def dims(self):
        out = (C.c_int64 * (self.n + 1))()
        check(self.ctx.handle, lib.qtt_mpo_dims(self.handle, out))
        return list(out)

    def free(self):
        if getattr(self, "handle", None):
            lib.qtt_mpo_free(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def linkdims(cores):
    return [c.shape[2] for c in cores[:-1]]


def maxlinkdim(cores):
    return max(linkdims(cores) + [1])
