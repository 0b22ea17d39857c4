"""Host-side mirror of the plugin's ``Groth16<E>: ProofSystem``
(/root/reference/plugins/arkworks/src/groth16.rs:399-467) for the device prover.

* ``Groth16.compile(r1cs, rng)``  ~ ``ProofSystem::compile`` (groth16.rs:428-443): circuit-specific
  setup.  The reference samples the trapdoor inside ``ark_groth16::generate_random_parameters``;
  here the same query vectors are computed from an explicit trapdoor (tau, alpha, beta, gamma, delta)
  with the device doing the heavy parts (one inverse NTT for the Lagrange basis at tau, transposed
  SpMVs for a_j/b_j/c_j, fixed-base scalar multiplications).  Setup is not the hot path.
* ``Groth16.prove(pk, z, rng)``   ~ ``ProofSystem::prove`` (groth16.rs:446-457): draws r, s like
  ``create_random_proof`` and runs ``ozl_groth16_prove`` (device witness map + 5 MSMs).
* ``verify`` stays with arkworks on the CPU in the reference (3 pairings, milliseconds;
  groth16.rs:460-466) and is not accelerated here; tests check proofs with the oracle.
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass
from typing import List, Optional, Tuple

import numpy as np

from . import _lib
from .circuits.r1cs import Csr, R1CS
from .context import Context

PAIRINGS = {
    "bn254": dict(id=0, g1=_lib.BN254_G1, g2=_lib.BN254_G2, fr=_lib.BN254_FR,
                  r=21888242871839275222246405745257275088548364400416034343698204186575808495617),
    "bls12_381": dict(id=1, g1=_lib.BLS12_381_G1, g2=_lib.BLS12_381_G2, fr=_lib.BLS12_381_FR,
                      r=0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001),
}
_R256 = 1 << 256


def ints_to_limbs(vals, modulus: Optional[int] = None, mont: bool = False) -> np.ndarray:
    """Canonical ints -> (n, 4) uint64; with mont=True the Montgomery residue v * 2^256 mod p."""
    n = len(vals)
    buf = bytearray(32 * n)
    if mont:
        for i, v in enumerate(vals):
            buf[32 * i:32 * i + 32] = ((v << 256) % modulus).to_bytes(32, "little")
    else:
        for i, v in enumerate(vals):
            buf[32 * i:32 * i + 32] = int(v).to_bytes(32, "little")
    return np.frombuffer(bytes(buf), dtype=np.uint64).reshape(n, 4).copy()


def limbs_to_ints(arr: np.ndarray, modulus: Optional[int] = None, mont: bool = False) -> List[int]:
    raw = np.ascontiguousarray(arr, dtype=np.uint64).tobytes()
    vals = [int.from_bytes(raw[32 * i:32 * i + 32], "little") for i in range(len(raw) // 32)]
    if mont:
        rinv = pow(_R256, -1, modulus)
        vals = [(v * rinv) % modulus for v in vals]
    return vals


def _csr_struct(m: Csr):
    rp = np.ascontiguousarray(m.row_ptr, dtype=np.uint32)
    ci = np.ascontiguousarray(m.col_idx, dtype=np.uint32)
    cf = np.ascontiguousarray(m.coef_idx, dtype=np.uint32)
    s = _lib.Csr(m.n_rows, rp.ctypes.data, ci.ctypes.data, cf.ctypes.data)
    return s, (rp, ci, cf)


@dataclass
class Trapdoor:
    tau: int
    alpha: int
    beta: int
    gamma: int
    delta: int


@dataclass
class VerifyingData:
    """What a verifier needs, kept as scalars because the trapdoor is known in this harness:
    ic[j] = (beta a_j + alpha b_j + c_j) / gamma for the instance variables."""
    pairing: str
    trapdoor: Trapdoor
    ic: List[int]


@dataclass
class Proof:
    a: np.ndarray   # G1 affine x||y  (Montgomery limbs)
    b: np.ndarray   # G2 affine
    c: np.ndarray   # G1 affine

    def to_bytes(self, pairing: str) -> bytes:
        """``proof_as_bytes`` (groth16.rs:98-107): ark's compressed a || b || c."""
        from . import serialize
        return serialize.proof_limbs_as_bytes(pairing, self.a, self.b, self.c)


class ProvingContext:
    """``ProvingContext<E>(ProvingKey<E>)`` resident on the device."""

    def __init__(self, ctx: Context, pairing: str, handle: int, r1cs: R1CS, domain_size: int, queries=None):
        self.ctx, self.pairing, self.handle, self.r1cs, self.domain_size = ctx, pairing, handle, r1cs, domain_size
        self.queries = queries   # host copies of the query scalars (tests only; None when not retained)

    def free(self):
        if self.handle:
            self.ctx._lib.ozl_groth16_pk_destroy(self.ctx._h, self.handle)
            self.handle = 0


def fr_spmv(ctx: Context, field: int, m: Csr, coef_table_mont: np.ndarray, x_mont: np.ndarray) -> np.ndarray:
    s, keep = _csr_struct(m)
    x_mont = np.ascontiguousarray(x_mont, dtype=np.uint64)
    y = np.zeros((m.n_rows, 4), dtype=np.uint64)
    ctx._check(ctx._lib.ozl_fr_spmv(ctx._h, field, ctypes.byref(s), coef_table_mont.ctypes.data, coef_table_mont.shape[0],
                                    x_mont.ctypes.data, x_mont.shape[0], y.ctypes.data), "ozl_fr_spmv")
    del keep
    return y


def fixed_base_mul(ctx: Context, curve: int, scalars: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """[k_j]G for canonical (n, 4) uint64 scalars -> (affine (n, 2L) uint64, identity bitset)."""
    scalars = np.ascontiguousarray(scalars, dtype=np.uint64)
    n = scalars.shape[0]
    limbs = ctx._lib.ozl_curve_coord_limbs(curve)
    out = np.zeros((n, 2 * limbs), dtype=np.uint64)
    flags = np.zeros(n, dtype=np.uint8)
    ctx._check(ctx._lib.ozl_fixed_base_mul(ctx._h, curve, scalars.ctypes.data, n, out.ctypes.data, flags.ctypes.data),
               "ozl_fixed_base_mul")
    return out, np.packbits(flags, bitorder="little")


class Groth16:
    """Associated-function style like the reference's zero-sized ``Groth16<E>(PhantomData)``."""

    @staticmethod
    def compile(ctx: Context, pairing: str, r1cs: R1CS, trapdoor: Trapdoor, keep_queries: bool = False,
                precompute: int = 4):
        """Known-trapdoor circuit-specific setup -> (ProvingContext, VerifyingData)."""
        P = PAIRINGS[pairing]
        p = P["r"]
        assert r1cs.modulus == p
        nc, ni, m = r1cs.n_constraints, r1cs.n_instance, r1cs.n_vars
        n = 1
        while n < nc + ni:
            n <<= 1
        t = trapdoor
        # Lagrange basis at tau: L = ifft(1, tau, tau^2, ...)   (L_i(tau) = (1/n) sum_k (tau w^-i)^k)
        pw, cur = [], 1
        for _ in range(n):
            pw.append(cur)
            cur = (cur * t.tau) % p
        L_m = ints_to_limbs(pw, p, mont=True)
        ctx.ntt(P["fr"], L_m, inverse=True)
        coef_m = ints_to_limbs(r1cs.coef_table, p, mont=True)
        abc = []
        for M in (r1cs.A, r1cs.B, r1cs.C):
            y = fr_spmv(ctx, P["fr"], M.transpose(m), coef_m, L_m[:nc])
            abc.append(limbs_to_ints(y, p, mont=True))
        a, b, c = abc
        L = limbs_to_ints(L_m[nc:nc + ni], p, mont=True)
        for j in range(ni):                       # ark's input-consistency rows: a[nc + j] = z_j
            a[j] = (a[j] + L[j]) % p
        zt = (pow(t.tau, n, p) - 1) % p
        dinv, ginv = pow(t.delta, -1, p), pow(t.gamma, -1, p)
        k = [(t.beta * a[j] + t.alpha * b[j] + c[j]) % p for j in range(m)]
        ic = [(k[j] * ginv) % p for j in range(ni)]
        lq = [(k[j] * dinv) % p for j in range(ni, m)]
        hq, cur = [], (zt * dinv) % p
        for _ in range(n - 1):
            hq.append(cur)
            cur = (cur * t.tau) % p

        def upload(curve, scal):
            pts, inf = fixed_base_mul(ctx, curve, ints_to_limbs(scal))
            hb = ctx.upload_bases(curve, pts, inf)
            return hb.precompute(precompute) if precompute > 1 else hb

        h_a, h_b1, h_b2 = upload(P["g1"], a), upload(P["g1"], b), upload(P["g2"], b)
        h_h, h_l = upload(P["g1"], hq), upload(P["g1"], lq)
        consts1, _ = fixed_base_mul(ctx, P["g1"], ints_to_limbs([t.alpha, t.beta, t.delta]))
        consts2, _ = fixed_base_mul(ctx, P["g2"], ints_to_limbs([t.beta, t.delta]))
        sa, ka = _csr_struct(r1cs.A)
        sb, kb = _csr_struct(r1cs.B)
        sc, kc = _csr_struct(r1cs.C)
        handle = ctypes.c_uint32(0)
        ctx._check(ctx._lib.ozl_groth16_pk_create(
            ctx._h, P["id"], nc, ni, m, ctypes.byref(sa), ctypes.byref(sb), ctypes.byref(sc), coef_m.ctypes.data,
            coef_m.shape[0], h_a.handle, h_b1.handle, h_b2.handle, h_h.handle, h_l.handle,
            consts1[0].ctypes.data, consts1[1].ctypes.data, consts1[2].ctypes.data, consts2[0].ctypes.data,
            consts2[1].ctypes.data, ctypes.byref(handle)), "ozl_groth16_pk_create")
        for hb in (h_a, h_b1, h_b2, h_h, h_l):
            hb.handle = 0                          # ownership moved into the pk
        del ka, kb, kc
        queries = dict(a=a, b=b, c=c, h=hq, l=lq) if keep_queries else None
        return ProvingContext(ctx, pairing, handle.value, r1cs, n, queries), VerifyingData(pairing, t, ic)

    @staticmethod
    def prove_with_randomness(pk: ProvingContext, z_mont: np.ndarray, r: int, s: int, want_h: bool = False):
        """``create_proof(circuit, pk, r, s)``; z_mont = full assignment, (n_vars, 4) uint64 Montgomery."""
        ctx = pk.ctx
        z_mont = np.ascontiguousarray(z_mont, dtype=np.uint64)
        if z_mont.shape != (pk.r1cs.n_vars, 4):
            raise _lib.OzlError(1, "groth16.prove", "assignment has the wrong shape")
        l1 = ctx._lib.ozl_curve_coord_limbs(PAIRINGS[pk.pairing]["g1"])
        l2 = ctx._lib.ozl_curve_coord_limbs(PAIRINGS[pk.pairing]["g2"])
        pa, pb, pc = np.zeros(2 * l1, dtype=np.uint64), np.zeros(2 * l2, dtype=np.uint64), np.zeros(2 * l1, dtype=np.uint64)
        rr, ss = ints_to_limbs([r]), ints_to_limbs([s])
        h = np.zeros((pk.domain_size, 4), dtype=np.uint64) if want_h else None
        ctx._check(ctx._lib.ozl_groth16_prove(ctx._h, pk.handle, z_mont.ctypes.data, rr.ctypes.data, ss.ctypes.data,
                                              pa.ctypes.data, pb.ctypes.data, pc.ctypes.data,
                                              h.ctypes.data if want_h else None), "ozl_groth16_prove")
        proof = Proof(pa, pb, pc)
        return (proof, h) if want_h else proof

    @staticmethod
    def prove(pk: ProvingContext, z_mont: np.ndarray, rng) -> Proof:
        """``ProofSystem::prove``: r, s drawn from the caller's rng first, as ark's
        ``create_random_proof`` does (two ``Fr::rand`` calls before any MSM)."""
        p = PAIRINGS[pk.pairing]["r"]
        r = int(rng.integers(0, 1 << 62)) * (1 << 192) % p if hasattr(rng, "integers") else rng.randrange(p)
        s = int(rng.integers(0, 1 << 62)) * (1 << 190) % p if hasattr(rng, "integers") else rng.randrange(p)
        return Groth16.prove_with_randomness(pk, z_mont, r, s)
