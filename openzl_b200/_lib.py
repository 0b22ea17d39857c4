"""ctypes binding of ``libozl_b200.so`` (the C ABI declared in ``include/ozl.h``).

There is deliberately no fallback: if the shared library is missing or cannot be loaded the
import of any compute entry point raises ``OzlLibraryError``.
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libozl_b200.so")

OZL_OK = 0
STATUS_NAMES = {0: "OZL_OK", 1: "OZL_ERR_ARG", 2: "OZL_ERR_CUDA", 3: "OZL_ERR_NO_DEVICE", 4: "OZL_ERR_OOM",
                5: "OZL_ERR_HANDLE", 6: "OZL_ERR_DOMAIN", 7: "OZL_ERR_NCCL"}

# enum ozl_curve / ozl_field (include/ozl.h)
BLS12_381_G1, BLS12_381_G2, BN254_G1, BN254_G2 = 0, 1, 2, 3
BN254_FR, BLS12_381_FR = 0, 1
CURVE_IDS = {"bls12_381_g1": 0, "bls12_381_g2": 1, "bn254_g1": 2, "bn254_g2": 3}
FIELD_IDS = {"bn254_fr": 0, "bls12_381_fr": 1}

# every symbol include/ozl.h declares (checked by tests/test_abi.py)
EXPORTS = [
    "ozl_version", "ozl_strerror", "ozl_last_error", "ozl_ctx_create", "ozl_ctx_destroy", "ozl_ctx_set_stream", "ozl_ctx_use_own_stream",
    "ozl_ctx_get_stream", "ozl_ctx_synchronize", "ozl_curve_coord_limbs", "ozl_msm_bases_upload",
    "ozl_msm_bases_upload_device", "ozl_msm_bases_generate", "ozl_msm_bases_precompute", "ozl_msm_bases_download", "ozl_msm_bases_free",
    "ozl_msm", "ozl_msm_submit", "ozl_msm_device_async", "ozl_msm_set_window_bits", "ozl_msm_set_batch_affine", "ozl_msm_get_window_bits", "ozl_msm_bases_info", "ozl_jacobian_sum",
    "ozl_jacobian_to_affine", "ozl_ntt", "ozl_ntt_device_async", "ozl_ctx_enable_timing",
    "ozl_ctx_get_stage_times", "ozl_ctx_get_stage_spans", "ozl_ctx_launch_count", "ozl_bench_field_mul", "ozl_fr_spmv", "ozl_fr_poseidon_permute", "ozl_fixed_base_mul",
    "ozl_groth16_pk_create", "ozl_groth16_pk_destroy", "ozl_groth16_prove", "ozl_groth16_domain_size",
    "ozl_comm_unique_id", "ozl_comm_create", "ozl_comm_destroy", "ozl_msm_sharded", "ozl_msm_sharded_device_async",
    "ozl_comm_allgather_sum_async",
]


class OzlLibraryError(RuntimeError):
    """libozl_b200.so is missing / unloadable: the CUDA extension was not built."""


class OzlError(RuntimeError):
    """A C-ABI call returned a non-zero status (mirrors the plugin's opaque ``Error``,
    /root/reference/plugins/arkworks/src/groth16.rs:35-45, plus a status code for debugging)."""

    def __init__(self, status: int, where: str, detail: str = ""):
        self.status = status
        super().__init__(f"{where}: {STATUS_NAMES.get(status, status)}" + (f" ({detail})" if detail else ""))


class Csr(ctypes.Structure):
    _fields_ = [("n_rows", ctypes.c_uint32), ("row_ptr", ctypes.c_void_p), ("col_idx", ctypes.c_void_p),
                ("coef_idx", ctypes.c_void_p)]


class StageTime(ctypes.Structure):
    _fields_ = [("name", ctypes.c_char * 32), ("ms", ctypes.c_float), ("launches", ctypes.c_int)]


class StageSpan(ctypes.Structure):
    _fields_ = [("name", ctypes.c_char * 32), ("start_ms", ctypes.c_float), ("end_ms", ctypes.c_float), ("launches", ctypes.c_int)]


_lib = None


def load() -> ctypes.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise OzlLibraryError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C openzl_b200/csrc`. There is no CPU fallback.")
    try:
        lib = ctypes.CDLL(LIB_PATH)
    except OSError as exc:  # pragma: no cover - depends on the host
        raise OzlLibraryError(f"cannot load {LIB_PATH}: {exc}") from exc
    vp, sz, u32p = ctypes.c_void_p, ctypes.c_size_t, ctypes.POINTER(ctypes.c_uint32)
    lib.ozl_version.restype = ctypes.c_int
    lib.ozl_strerror.restype = ctypes.c_char_p
    lib.ozl_strerror.argtypes = [ctypes.c_int]
    lib.ozl_last_error.restype = ctypes.c_char_p
    lib.ozl_last_error.argtypes = [vp]
    lib.ozl_ctx_create.argtypes = [ctypes.c_int, ctypes.POINTER(vp)]
    lib.ozl_ctx_destroy.argtypes = [vp]
    lib.ozl_ctx_destroy.restype = None
    lib.ozl_ctx_set_stream.argtypes = [vp, vp]
    lib.ozl_ctx_use_own_stream.argtypes = [vp]
    lib.ozl_ctx_get_stream.argtypes = [vp]
    lib.ozl_ctx_get_stream.restype = vp
    lib.ozl_ctx_synchronize.argtypes = [vp]
    lib.ozl_curve_coord_limbs.argtypes = [ctypes.c_int]
    lib.ozl_msm_bases_upload.argtypes = [vp, ctypes.c_int, vp, vp, sz, u32p]
    lib.ozl_msm_bases_upload_device.argtypes = [vp, ctypes.c_int, vp, vp, sz, u32p]
    lib.ozl_msm_bases_generate.argtypes = [vp, ctypes.c_int, ctypes.c_uint64, sz, u32p]
    lib.ozl_msm_bases_precompute.argtypes = [vp, ctypes.c_uint32, ctypes.c_int]
    lib.ozl_msm_bases_download.argtypes = [vp, ctypes.c_uint32, sz, sz, vp]
    lib.ozl_msm_bases_free.argtypes = [vp, ctypes.c_uint32]
    lib.ozl_msm.argtypes = [vp, ctypes.c_uint32, vp, sz, vp]
    lib.ozl_msm_submit.argtypes = [vp, ctypes.c_uint32, vp, sz, vp]
    lib.ozl_msm_device_async.argtypes = [vp, ctypes.c_uint32, vp, sz, vp]
    lib.ozl_msm_set_window_bits.argtypes = [vp, ctypes.c_int]
    lib.ozl_comm_unique_id.argtypes = [vp]
    lib.ozl_comm_create.argtypes = [vp, vp, ctypes.c_int, ctypes.c_int, ctypes.POINTER(vp)]
    lib.ozl_comm_destroy.argtypes = [vp]
    lib.ozl_msm_sharded.argtypes = [vp, vp, ctypes.c_uint32, vp, sz, vp]
    lib.ozl_msm_sharded_device_async.argtypes = [vp, vp, ctypes.c_uint32, vp, sz, vp]
    lib.ozl_comm_allgather_sum_async.argtypes = [vp, vp, ctypes.c_int, vp, vp]
    lib.ozl_msm_set_batch_affine.argtypes = [vp, ctypes.c_int]
    lib.ozl_msm_get_window_bits.argtypes = [vp, ctypes.c_int, sz]
    ip = ctypes.POINTER(ctypes.c_int)
    lib.ozl_msm_bases_info.argtypes = [vp, ctypes.c_uint32, sz, ip, ip, ip, ip]
    lib.ozl_jacobian_sum.argtypes = [vp, ctypes.c_int, vp, sz, vp]
    lib.ozl_jacobian_to_affine.argtypes = [vp, ctypes.c_int, vp, vp, ctypes.POINTER(ctypes.c_int)]
    lib.ozl_ntt.argtypes = [vp, ctypes.c_int, vp, ctypes.c_uint32, ctypes.c_int, ctypes.c_int]
    lib.ozl_ntt_device_async.argtypes = [vp, ctypes.c_int, vp, ctypes.c_uint32, ctypes.c_int, ctypes.c_int]
    lib.ozl_ctx_enable_timing.argtypes = [vp, ctypes.c_int]
    lib.ozl_ctx_get_stage_times.argtypes = [vp, ctypes.POINTER(StageTime), ctypes.c_int]
    lib.ozl_ctx_get_stage_spans.argtypes = [vp, ctypes.POINTER(StageSpan), ctypes.c_int]
    lib.ozl_ctx_launch_count.argtypes = [vp]
    lib.ozl_ctx_launch_count.restype = ctypes.c_uint64
    lib.ozl_bench_field_mul.argtypes = [vp, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_double)]
    csrp = ctypes.POINTER(Csr)
    u32 = ctypes.c_uint32
    lib.ozl_fr_spmv.argtypes = [vp, ctypes.c_int, csrp, vp, u32, vp, u32, vp]
    lib.ozl_fr_poseidon_permute.argtypes = [vp, ctypes.c_int, vp, sz, u32, u32, u32, vp, vp]
    lib.ozl_fixed_base_mul.argtypes = [vp, ctypes.c_int, vp, sz, vp, vp]
    lib.ozl_groth16_pk_create.argtypes = [vp, ctypes.c_int, u32, u32, u32, csrp, csrp, csrp, vp, u32, u32, u32, u32, u32, u32,
                                          vp, vp, vp, vp, vp, u32p]
    lib.ozl_groth16_pk_destroy.argtypes = [vp, u32]
    lib.ozl_groth16_prove.argtypes = [vp, u32, vp, vp, vp, vp, vp, vp, vp]
    lib.ozl_groth16_domain_size.argtypes = [vp, u32, u32p]
    _lib = lib
    return lib
