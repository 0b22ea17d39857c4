"""Host-side mirror of the plugin's ``Groth16<E>: ProofSystem``
(/root/reference/plugins/arkworks/src/groth16.rs:399-467) for the device prover.

* ``Groth16.compile(r1cs, rng)``  ~ ``ProofSystem::compile`` (groth16.rs:428-443): circuit-specific
  setup.  The reference samples the trapdoor inside ``ark_groth16::generate_random_parameters``;
  here the same query vectors are computed from an explicit trapdoor (tau, alpha, beta, gamma, delta)
  with the device doing the heavy parts (one inverse NTT for the Lagrange basis at tau, transposed
  SpMVs for a_j/b_j/c_j, fixed-base scalar multiplications).  Setup is not the hot path.
* ``Groth16.prove(pk, z, rng)``   ~ ``ProofSystem::prove`` (groth16.rs:446-457): draws r, s like
  ``create_random_proof`` and runs ``ozl_groth16_prove`` (device witness map + 5 MSMs).
* ``Groth16.verify(vk, input, proof)`` ~ ``ProofSystem::verify`` (groth16.rs:459-466): host-side, like the
  reference (three pairings, milliseconds; SURVEY.md row a-6 keeps it off the GPU).  The pairing is
  ``openzl_b200.pairing`` (optimal ate over the Fq2/Fq6/Fq12 tower).
* ``ProvingContext.encode`` / ``Groth16.proving_context_from_bytes`` ~ ``ProvingContext::{encode, decode}``
  (groth16.rs:142-179): ark's uncompressed, unchecked ``ProvingKey`` bytes <-> device-resident MSM bases.
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass
from typing import List, Optional, Tuple

import numpy as np

from . import _lib
from .circuits.r1cs import Csr, R1CS
from .context import Bases, Context

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
    """``VerifyingContext`` (groth16.rs:183-186): the verifying key as affine points in the C ABI's
    layout (Montgomery limbs), ``gamma_abc_g1[j] = [(beta a_j + alpha b_j + c_j) / gamma] G1`` for the
    instance variables.  A key made by the known-trapdoor ``compile`` also records the trapdoor and the
    ``ic`` scalars (harness only: tests compare proofs with their discrete logs); a key decoded from
    bytes has neither."""
    pairing: str
    trapdoor: Optional[Trapdoor]
    ic: Optional[List[int]]
    alpha_g1: Optional[np.ndarray] = None
    beta_g2: Optional[np.ndarray] = None
    gamma_g2: Optional[np.ndarray] = None
    delta_g2: Optional[np.ndarray] = None
    gamma_abc_g1: Optional[np.ndarray] = None      # (n_instance, 2 * limbs)

    def to_serializable(self):
        from . import serialize as ser
        g1, g2 = ser.PAIRING_GROUPS[self.pairing]
        return ser.VerifyingKey(ser.limbs_to_point(g1, self.alpha_g1), ser.limbs_to_point(g2, self.beta_g2),
                                ser.limbs_to_point(g2, self.gamma_g2), ser.limbs_to_point(g2, self.delta_g2),
                                ser.limbs_to_points(g1, self.gamma_abc_g1, _zero_rows_mask(self.gamma_abc_g1)))

    def to_bytes(self) -> bytes:
        """``VerifyingContext::encode``-style bytes: ark's compressed ``VerifyingKey``."""
        from . import serialize as ser
        return ser.vk_to_bytes(self.pairing, self.to_serializable())

    @staticmethod
    def from_serializable(pairing: str, vk) -> "VerifyingData":
        from . import serialize as ser
        g1, g2 = ser.PAIRING_GROUPS[pairing]
        abc, _ = ser.points_to_limbs(g1, vk.gamma_abc_g1)
        return VerifyingData(pairing, None, None, ser.point_to_limbs(g1, vk.alpha_g1), ser.point_to_limbs(g2, vk.beta_g2),
                             ser.point_to_limbs(g2, vk.gamma_g2), ser.point_to_limbs(g2, vk.delta_g2), abc)


def _zero_rows_mask(arr: np.ndarray) -> np.ndarray:
    """Infinity bitset of an array of affine points in which (0, 0) encodes the identity."""
    return np.packbits(~np.asarray(arr).reshape(len(arr), -1).any(axis=1), bitorder="little")


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

    def __init__(self, ctx: Context, pairing: str, handle: int, r1cs: R1CS, domain_size: int, queries=None,
                 query_handles=None, vk: Optional[VerifyingData] = None, beta_g1=None, delta_g1=None):
        self.ctx, self.pairing, self.handle, self.r1cs, self.domain_size = ctx, pairing, handle, r1cs, domain_size
        self.queries = queries   # host copies of the query scalars (tests only; None when not retained)
        self.query_handles = query_handles or {}   # bases handles owned by the device pk (kept for `encode`)
        self.vk, self.beta_g1, self.delta_g1 = vk, beta_g1, delta_g1

    def encode(self) -> bytes:
        """``ProvingContext::encode`` (groth16.rs:163-179): ark's ``serialize_unchecked`` of the
        ``ProvingKey`` -- the query vectors are read back from the device."""
        from . import serialize as ser
        g1, g2 = ser.PAIRING_GROUPS[self.pairing]
        P = PAIRINGS[self.pairing]

        def query(name, curve, g):
            h, n = self.query_handles[name]
            arr = Bases(self.ctx, h, curve, n).download()
            return ser.limbs_to_points(g, arr, _zero_rows_mask(arr))

        pk = ser.ProvingKey(self.vk.to_serializable(), ser.limbs_to_point(g1, self.beta_g1), ser.limbs_to_point(g1, self.delta_g1),
                            query("a", P["g1"], g1), query("b1", P["g1"], g1), query("b2", P["g2"], g2),
                            query("h", P["g1"], g1), query("l", P["g1"], g1))
        return ser.proving_key_to_bytes(self.pairing, pk)

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
        gamma2, _ = fixed_base_mul(ctx, P["g2"], ints_to_limbs([t.gamma]))
        abc1, _ = fixed_base_mul(ctx, P["g1"], ints_to_limbs(ic))
        vk = VerifyingData(pairing, t, ic, consts1[0].copy(), consts2[0].copy(), gamma2[0].copy(), consts2[1].copy(), abc1)
        handles = dict(a=h_a, b1=h_b1, b2=h_b2, h=h_h, l=h_l)
        pkc = Groth16._create_pk(ctx, pairing, r1cs, n, handles, consts1[0], consts1[1], consts1[2], consts2[0], consts2[1], vk)
        pkc.queries = dict(a=a, b=b, c=c, h=hq, l=lq) if keep_queries else None
        return pkc, vk

    @staticmethod
    def _create_pk(ctx: Context, pairing: str, r1cs: R1CS, n: int, handles, alpha1, beta1, delta1, beta2, delta2, vk):
        P = PAIRINGS[pairing]
        coef_m = ints_to_limbs(r1cs.coef_table, P["r"], mont=True)
        sa, ka = _csr_struct(r1cs.A)
        sb, kb = _csr_struct(r1cs.B)
        sc, kc = _csr_struct(r1cs.C)
        alpha1, beta1, delta1, beta2, delta2 = [np.ascontiguousarray(v, dtype=np.uint64) for v in (alpha1, beta1, delta1, beta2, delta2)]
        handle = ctypes.c_uint32(0)
        ctx._check(ctx._lib.ozl_groth16_pk_create(
            ctx._h, P["id"], r1cs.n_constraints, r1cs.n_instance, r1cs.n_vars, ctypes.byref(sa), ctypes.byref(sb), ctypes.byref(sc),
            coef_m.ctypes.data, coef_m.shape[0], handles["a"].handle, handles["b1"].handle, handles["b2"].handle,
            handles["h"].handle, handles["l"].handle, alpha1.ctypes.data, beta1.ctypes.data, delta1.ctypes.data,
            beta2.ctypes.data, delta2.ctypes.data, ctypes.byref(handle)), "ozl_groth16_pk_create")
        qh = {}
        for name, hb in handles.items():
            qh[name] = (hb.handle, hb.n)
            hb.handle = 0                          # ownership moved into the pk
        del ka, kb, kc
        return ProvingContext(ctx, pairing, handle.value, r1cs, n, None, qh, vk, beta1.copy(), delta1.copy())

    @staticmethod
    def proving_context_from_bytes(ctx: Context, pairing: str, raw: bytes, r1cs: R1CS, precompute: int = 4):
        """``ProvingContext::decode`` (groth16.rs:142-160) onto the device: ark's unchecked, uncompressed
        ``ProvingKey`` bytes -> five MSM bases uploads (`ozl_msm_bases_upload`, with the infinity bitsets) ->
        a device proving key for the circuit `r1cs`.  Returns (ProvingContext, VerifyingData)."""
        from . import serialize as ser
        P = PAIRINGS[pairing]
        g1, g2 = ser.PAIRING_GROUPS[pairing]
        pk = ser.proving_key_from_bytes(pairing, raw)
        nc, ni, m = r1cs.n_constraints, r1cs.n_instance, r1cs.n_vars
        n = 1
        while n < nc + ni:
            n <<= 1
        if (len(pk.a_query), len(pk.b_g1_query), len(pk.b_g2_query), len(pk.h_query), len(pk.l_query)) != (m, m, m, n - 1, m - ni) \
                or len(pk.vk.gamma_abc_g1) != ni:
            raise _lib.OzlError(1, "proving_context_from_bytes", "the key does not belong to this circuit (query lengths differ)")

        def upload(curve, g, pts):
            arr, mask = ser.points_to_limbs(g, pts)
            hb = ctx.upload_bases(curve, arr, mask)
            return hb.precompute(precompute) if precompute > 1 else hb

        handles = dict(a=upload(P["g1"], g1, pk.a_query), b1=upload(P["g1"], g1, pk.b_g1_query), b2=upload(P["g2"], g2, pk.b_g2_query),
                       h=upload(P["g1"], g1, pk.h_query), l=upload(P["g1"], g1, pk.l_query))
        vk = VerifyingData.from_serializable(pairing, pk.vk)
        pkc = Groth16._create_pk(ctx, pairing, r1cs, n, handles, vk.alpha_g1, ser.point_to_limbs(g1, pk.beta_g1),
                                 ser.point_to_limbs(g1, pk.delta_g1), vk.beta_g2, vk.delta_g2, vk)
        return pkc, vk

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
    def prove(pk: ProvingContext, z_mont: np.ndarray, rng=None) -> Proof:
        """``ProofSystem::prove``: r, s drawn first, as ark's ``create_random_proof`` does (two
        ``Fr::rand`` calls before any MSM), each UNIFORM in [0, r) by ark's mask-and-reject over 256 fresh
        bits.  The reference demands ``CryptoRng`` (constraint.rs:73-79): low-entropy or predictable r, s
        void zero-knowledge.  `rng` = None uses the operating system's CSPRNG (``secrets``); otherwise it
        must be a cryptographic generator exposing ``randbytes(n)`` or ``getrandbits(k)`` (e.g.
        ``random.SystemRandom``) -- numpy Generators and ``random.Random`` are refused."""
        p = PAIRINGS[pk.pairing]["r"]
        r = _uniform_scalar(p, rng)
        s = _uniform_scalar(p, rng)
        return Groth16.prove_with_randomness(pk, z_mont, r, s)

    @staticmethod
    def verify(vk: VerifyingData, public_inputs, proof: Proof) -> bool:
        """``ProofSystem::verify`` (groth16.rs:459-466 -> ``verify_with_processed_vk``):
        e(A, B) = e(alpha, beta) e(sum_j x_j gamma_abc_j, gamma) e(C, delta), x_0 = 1, on the host."""
        from . import pairing as pr
        return pr.groth16_verify(vk.pairing, vk.alpha_g1, vk.beta_g2, vk.gamma_g2, vk.delta_g2, vk.gamma_abc_g1,
                                 [int(x) for x in public_inputs], proof.a, proof.b, proof.c)


def _uniform_scalar(p: int, rng=None) -> int:
    """ark ``UniformRand`` for a prime field: sample the limbs, mask the bits above the modulus, reject
    values >= p.  Only cryptographic sources are accepted."""
    import random as _random
    import secrets
    bits = p.bit_length()
    if rng is None:
        draw = lambda: secrets.randbits(bits)
    elif isinstance(rng, _random.SystemRandom):
        draw = lambda: rng.getrandbits(bits)
    elif isinstance(rng, _random.Random) or hasattr(rng, "integers") or hasattr(rng, "bit_generator"):
        raise TypeError("Groth16.prove needs a cryptographic rng (None = OS CSPRNG, or random.SystemRandom); "
                        "use prove_with_randomness for reproducible tests")
    elif hasattr(rng, "randbytes"):
        draw = lambda: int.from_bytes(rng.randbytes(32), "little") & ((1 << bits) - 1)
    elif hasattr(rng, "getrandbits"):
        draw = lambda: rng.getrandbits(bits)
    else:
        raise TypeError("rng must provide randbytes(n) or getrandbits(k)")
    while True:
        v = draw()
        if v < p:
            return v
