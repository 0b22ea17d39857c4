"""ctypes binding of the C++ oracle (``oracle/c/ozl_oracle.cpp``).  Test infrastructure only."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libozl_oracle.so")

CURVE_IDS = {"bls12_381_g1": 0, "bls12_381_g2": 1, "bn254_g1": 2, "bn254_g2": 3}
NTT_FIELD_IDS = {"bn254_fr": 0, "bls12_381_fr": 1}
FP_FIELD_IDS = {"bls12_381_fq": 0, "bls12_381_fr": 1, "bn254_fq": 2, "bn254_fr": 3}
# u64 limbs per coordinate
COORD_LIMBS = {"bls12_381_g1": 6, "bls12_381_g2": 12, "bn254_g1": 4, "bn254_g2": 8}


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "c", "ozl_oracle.cpp")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        # -march=native: the .so is rebuilt on whatever box runs it if the CPU differs
        subprocess.check_call(["make", "-C", _HERE, "-B", "all"], stdout=subprocess.DEVNULL)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        try:
            build()
            _lib = ctypes.CDLL(_SO)
            _lib.oracle_hw_threads()
        except (OSError, subprocess.CalledProcessError):
            build(force=True)
            _lib = ctypes.CDLL(_SO)
        _lib.oracle_bases_seq.argtypes = [ctypes.c_int, ctypes.c_uint64, ctypes.c_size_t, ctypes.c_void_p]
        _lib.oracle_msm.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                    ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p]
        _lib.oracle_bases_from_dlogs.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]
        _lib.oracle_dot_mod_r.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]
        _lib.oracle_window_bits.argtypes = [ctypes.c_size_t]
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def hw_threads() -> int:
    return lib().oracle_hw_threads()


def window_bits(n: int) -> int:
    return lib().oracle_window_bits(n)


def bases_seq(curve: str, start: int, n: int) -> np.ndarray:
    """P_i = [start + i]G as an (n, 2*coord_limbs) uint64 array of Montgomery limbs."""
    out = np.zeros((n, 2 * COORD_LIMBS[curve]), dtype=np.uint64)
    assert lib().oracle_bases_seq(CURVE_IDS[curve], start, n, _p(out)) == 0
    return out


def bases_from_dlogs(curve: str, dlogs: np.ndarray) -> np.ndarray:
    dlogs = np.ascontiguousarray(dlogs, dtype=np.uint64)
    out = np.zeros((len(dlogs), 2 * COORD_LIMBS[curve]), dtype=np.uint64)
    assert lib().oracle_bases_from_dlogs(CURVE_IDS[curve], _p(dlogs), len(dlogs), _p(out)) == 0
    return out


def msm(curve: str, bases: np.ndarray, scalars: np.ndarray, inf: np.ndarray | None = None,
        threads: int = 1) -> np.ndarray:
    """ark-restatement MSM; returns Jacobian X||Y||Z Montgomery limbs (uint64)."""
    bases = np.ascontiguousarray(bases, dtype=np.uint64)
    scalars = np.ascontiguousarray(scalars, dtype=np.uint64)
    n = scalars.shape[0]
    out = np.zeros(3 * COORD_LIMBS[curve], dtype=np.uint64)
    assert lib().oracle_msm(CURVE_IDS[curve], _p(bases), _p(inf), _p(scalars), n, threads, _p(out)) == 0
    return out


def to_affine(curve: str, jac: np.ndarray):
    """Returns (affine limbs uint64[2*L], is_inf)."""
    jac = np.ascontiguousarray(jac, dtype=np.uint64)
    out = np.zeros(2 * COORD_LIMBS[curve], dtype=np.uint64)
    inf = ctypes.c_int(0)
    assert lib().oracle_to_affine(CURVE_IDS[curve], _p(jac), _p(out), ctypes.byref(inf)) == 0
    return out, bool(inf.value)


def gen_mul(curve: str, k: int) -> np.ndarray:
    kk = np.array([(k >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)], dtype=np.uint64)
    out = np.zeros(3 * COORD_LIMBS[curve], dtype=np.uint64)
    assert lib().oracle_gen_mul(CURVE_IDS[curve], _p(kk), _p(out)) == 0
    return out


def dot_mod_r(field: str, scalars: np.ndarray, dlogs: np.ndarray) -> int:
    scalars = np.ascontiguousarray(scalars, dtype=np.uint64)
    dlogs = np.ascontiguousarray(dlogs, dtype=np.uint64)
    out = np.zeros(4, dtype=np.uint64)
    assert lib().oracle_dot_mod_r(NTT_FIELD_IDS[field], _p(scalars), _p(dlogs), len(dlogs), _p(out)) == 0
    return sum(int(x) << (64 * i) for i, x in enumerate(out))


def ntt(field: str, data: np.ndarray, inverse: bool = False, coset: bool = False) -> np.ndarray:
    """In-order NTT of a (2^k, 4) uint64 Montgomery array; returns a new array."""
    a = np.array(data, dtype=np.uint64, order="C", copy=True)
    n = a.shape[0]
    log_n = n.bit_length() - 1
    assert 1 << log_n == n
    rc = lib().oracle_ntt(NTT_FIELD_IDS[field], _p(a), log_n, int(inverse), int(coset))
    if rc != 0:
        raise ValueError(f"oracle_ntt rc={rc}")
    return a


def fp_mul(field: str, a: np.ndarray, b: np.ndarray) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.uint64)
    b = np.ascontiguousarray(b, dtype=np.uint64)
    out = np.zeros_like(a)
    assert lib().oracle_fp_mul(FP_FIELD_IDS[field], _p(a), _p(b), _p(out)) == 0
    return out
