"""Device context and bases registry: the host-side objects behind the C ABI."""
from __future__ import annotations

import ctypes
from typing import Optional

import numpy as np

from . import _lib
from ._lib import OzlError


def _ptr(a) -> Optional[int]:
    """Pointer of a numpy array (host) or an int/None (device pointer passed through)."""
    if a is None:
        return None
    if isinstance(a, int):
        return a
    return a.ctypes.data


class Context:
    """One per (device, stream); wraps ``ozl_ctx``.  Thread-compatible, not thread-safe."""

    def __init__(self, device: int = 0):
        self._lib = _lib.load()
        h = ctypes.c_void_p()
        rc = self._lib.ozl_ctx_create(device, ctypes.byref(h))
        if rc != 0:
            raise OzlError(rc, "ozl_ctx_create", "a B200-class CUDA device is required; there is no CPU fallback")
        self._h = h
        self.device = device

    # -- plumbing ---------------------------------------------------------------------------
    def _check(self, rc: int, where: str):
        if rc != 0:
            detail = self._lib.ozl_last_error(self._h).decode(errors="replace")
            raise OzlError(rc, where, detail)

    def close(self):
        if getattr(self, "_h", None):
            self._lib.ozl_ctx_destroy(self._h)
            self._h = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def set_stream(self, cuda_stream_ptr: Optional[int]):
        """Run on the given cudaStream_t (0 / None = CUDA's legacy default stream)."""
        self._check(self._lib.ozl_ctx_set_stream(self._h, cuda_stream_ptr or None), "ozl_ctx_set_stream")

    def use_own_stream(self):
        self._check(self._lib.ozl_ctx_use_own_stream(self._h), "ozl_ctx_use_own_stream")

    def use_torch_stream(self, stream=None):
        """Run on a torch CUDA stream (default: the current one) so torch events bracket our kernels."""
        import torch
        s = stream if stream is not None else torch.cuda.current_stream(self.device)
        self.set_stream(s.cuda_stream)

    def synchronize(self):
        self._check(self._lib.ozl_ctx_synchronize(self._h), "ozl_ctx_synchronize")

    @property
    def launch_count(self) -> int:
        return int(self._lib.ozl_ctx_launch_count(self._h))

    def enable_timing(self, on: bool = True):
        self._check(self._lib.ozl_ctx_enable_timing(self._h, int(on)), "ozl_ctx_enable_timing")

    def stage_times(self):
        buf = (_lib.StageTime * 64)()
        k = self._lib.ozl_ctx_get_stage_times(self._h, buf, 64)
        if k < 0:
            raise OzlError(2, "ozl_ctx_get_stage_times")
        return [(buf[i].name.decode(), float(buf[i].ms), int(buf[i].launches)) for i in range(k)]

    def stage_spans(self):
        """Timeline of the most recent timed call: (name, start_ms, end_ms, launches), offsets from the first stage."""
        buf = (_lib.StageSpan * 96)()
        k = self._lib.ozl_ctx_get_stage_spans(self._h, buf, 96)
        if k < 0:
            raise OzlError(2, "ozl_ctx_get_stage_spans")
        return [(buf[i].name.decode(), float(buf[i].start_ms), float(buf[i].end_ms), int(buf[i].launches)) for i in range(k)]

    def bench_field_mul(self, field_id: int = 0, iters: int = 2000) -> float:
        out = ctypes.c_double(0.0)
        self._check(self._lib.ozl_bench_field_mul(self._h, field_id, iters, ctypes.byref(out)), "ozl_bench_field_mul")
        return out.value

    # -- MSM --------------------------------------------------------------------------------
    def set_window_bits(self, c: int):
        self._check(self._lib.ozl_msm_set_window_bits(self._h, c), "ozl_msm_set_window_bits")

    def set_batch_affine(self, levels: int):
        """EXPERIMENTAL batched-affine pair levels before the XYZZ accumulation (0 = off, -1 = default)."""
        self._check(self._lib.ozl_msm_set_batch_affine(self._h, levels), "ozl_msm_set_batch_affine")

    def window_bits(self, curve: int, n: int) -> int:
        return int(self._lib.ozl_msm_get_window_bits(self._h, curve, n))

    def upload_bases(self, curve: int, bases: np.ndarray, inf_mask: Optional[np.ndarray] = None) -> "Bases":
        limbs = self._lib.ozl_curve_coord_limbs(curve)
        if limbs <= 0:
            raise OzlError(1, "upload_bases", "unknown curve")
        bases = np.ascontiguousarray(bases, dtype=np.uint64)
        if bases.ndim != 2 or bases.shape[1] != 2 * limbs:
            raise OzlError(1, "upload_bases", f"expected shape (n, {2 * limbs}) uint64 Montgomery limbs")
        n = bases.shape[0]
        if inf_mask is not None:
            inf_mask = np.ascontiguousarray(inf_mask, dtype=np.uint8)
            if inf_mask.size < (n + 7) // 8:
                raise OzlError(1, "upload_bases", "inf_mask too short")
        h = ctypes.c_uint32(0)
        self._check(self._lib.ozl_msm_bases_upload(self._h, curve, _ptr(bases), _ptr(inf_mask), n, ctypes.byref(h)),
                    "ozl_msm_bases_upload")
        return Bases(self, h.value, curve, n)

    def upload_bases_device(self, curve: int, d_bases: int, n: int, d_inf_mask: Optional[int] = None) -> "Bases":
        h = ctypes.c_uint32(0)
        self._check(self._lib.ozl_msm_bases_upload_device(self._h, curve, d_bases, d_inf_mask, n, ctypes.byref(h)),
                    "ozl_msm_bases_upload_device")
        return Bases(self, h.value, curve, n)

    def generate_bases(self, curve: int, start: int, n: int) -> "Bases":
        """P_i = [start + i]G synthesized on the device (benchmarks, size-independent checks)."""
        h = ctypes.c_uint32(0)
        self._check(self._lib.ozl_msm_bases_generate(self._h, curve, start, n, ctypes.byref(h)), "ozl_msm_bases_generate")
        return Bases(self, h.value, curve, n)

    def jacobian_sum(self, curve: int, points: np.ndarray) -> np.ndarray:
        limbs = self._lib.ozl_curve_coord_limbs(curve)
        points = np.ascontiguousarray(points, dtype=np.uint64).reshape(-1, 3 * limbs)
        out = np.zeros(3 * limbs, dtype=np.uint64)
        self._check(self._lib.ozl_jacobian_sum(self._h, curve, _ptr(points), points.shape[0], _ptr(out)), "ozl_jacobian_sum")
        return out

    def jacobian_to_affine(self, curve: int, jac: np.ndarray):
        limbs = self._lib.ozl_curve_coord_limbs(curve)
        jac = np.ascontiguousarray(jac, dtype=np.uint64).reshape(3 * limbs)
        out = np.zeros(2 * limbs, dtype=np.uint64)
        flag = ctypes.c_int(0)
        self._check(self._lib.ozl_jacobian_to_affine(self._h, curve, _ptr(jac), _ptr(out), ctypes.byref(flag)),
                    "ozl_jacobian_to_affine")
        return out, bool(flag.value)

    # -- NTT --------------------------------------------------------------------------------
    def ntt(self, field: int, data: np.ndarray, inverse: bool = False, coset: bool = False) -> None:
        """In-place on a C-contiguous (2^k, 4) uint64 host array."""
        if data.dtype != np.uint64 or not data.flags["C_CONTIGUOUS"] or data.ndim != 2 or data.shape[1] != 4:
            raise OzlError(1, "ntt", "expected C-contiguous uint64 array of shape (2^k, 4)")
        n = data.shape[0]
        log_n = n.bit_length() - 1
        if n == 0 or (1 << log_n) != n:
            raise OzlError(1, "ntt", "length must be a power of two")
        self._check(self._lib.ozl_ntt(self._h, field, _ptr(data), log_n, int(inverse), int(coset)), "ozl_ntt")

    def ntt_device(self, field: int, d_data: int, log_n: int, inverse: bool = False, coset: bool = False) -> None:
        self._check(self._lib.ozl_ntt_device_async(self._h, field, d_data, log_n, int(inverse), int(coset)),
                    "ozl_ntt_device_async")


class Bases:
    """Device-resident MSM bases (the constant half of every Groth16 MSM)."""

    def __init__(self, ctx: Context, handle: int, curve: int, n: int):
        self.ctx, self.handle, self.curve, self.n = ctx, handle, curve, n
        self.coord_limbs = ctx._lib.ozl_curve_coord_limbs(curve)

    def free(self):
        if self.handle and self.ctx._h:
            self.ctx._lib.ozl_msm_bases_free(self.ctx._h, self.handle)
        self.handle = 0

    def precompute(self, factor: int) -> "Bases":
        """Store `factor` shifted copies of the bases (fewer bucket sets / shorter final Horner)."""
        self.ctx._check(self.ctx._lib.ozl_msm_bases_precompute(self.ctx._h, self.handle, factor), "ozl_msm_bases_precompute")
        return self

    def info(self, n: Optional[int] = None) -> dict:
        """Plan of an n-scalar MSM on this handle: window bits, windows, bucket sets, copies."""
        v = [ctypes.c_int(0) for _ in range(4)]
        self.ctx._check(self.ctx._lib.ozl_msm_bases_info(self.ctx._h, self.handle, self.n if n is None else n,
                                                         *[ctypes.byref(x) for x in v]), "ozl_msm_bases_info")
        return dict(c=v[0].value, windows=v[1].value, bucket_sets=v[2].value, factor=v[3].value)

    def download(self, first: int = 0, n: Optional[int] = None) -> np.ndarray:
        n = self.n - first if n is None else n
        out = np.zeros((n, 2 * self.coord_limbs), dtype=np.uint64)
        self.ctx._check(self.ctx._lib.ozl_msm_bases_download(self.ctx._h, self.handle, first, n, _ptr(out)),
                        "ozl_msm_bases_download")
        return out

    def msm(self, scalars: np.ndarray) -> np.ndarray:
        """Host scalars (n, 4) uint64 canonical -> Jacobian X||Y||Z (uint64 Montgomery limbs)."""
        scalars = np.ascontiguousarray(scalars, dtype=np.uint64)
        if scalars.ndim != 2 or scalars.shape[1] != 4:
            raise OzlError(1, "msm", "expected scalars of shape (n, 4) uint64 (ark BigInteger256)")
        n = scalars.shape[0]
        out = np.zeros(3 * self.coord_limbs, dtype=np.uint64)
        self.ctx._check(self.ctx._lib.ozl_msm(self.ctx._h, self.handle, _ptr(scalars), n, _ptr(out)), "ozl_msm")
        return out

    def msm_host_ptr(self, scalars_ptr: int, n: int, out: np.ndarray) -> None:
        self.ctx._check(self.ctx._lib.ozl_msm(self.ctx._h, self.handle, scalars_ptr, n, _ptr(out)), "ozl_msm")

    def msm_submit(self, scalars_ptr: int, n: int, out_ptr: int) -> None:
        """Pipelined host-buffer MSM (returns after enqueueing; `Context.synchronize()` completes it)."""
        self.ctx._check(self.ctx._lib.ozl_msm_submit(self.ctx._h, self.handle, scalars_ptr, n, out_ptr), "ozl_msm_submit")

    def msm_device(self, d_scalars: int, n: int, d_out: int) -> None:
        """Device pointers in/out; asynchronous on the context's stream."""
        self.ctx._check(self.ctx._lib.ozl_msm_device_async(self.ctx._h, self.handle, d_scalars, n, d_out),
                        "ozl_msm_device_async")
