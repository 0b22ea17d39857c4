"""Wire formats either side of the hot path: ark-serialize 0.3.0
``CanonicalSerialize`` for the values the plugin moves across its codec boundary.

What the reference does with these bytes:

* ``proof_as_bytes`` -- ``proof.serialize(&mut buffer)``, the *compressed* form
  (/root/reference/plugins/arkworks/src/groth16.rs:98-107); ``Proof: TryFrom<Vec<u8>>`` reads it back
  with ``CanonicalDeserialize::deserialize`` (groth16.rs:86-92).
* ``ProvingContext::encode`` -- ``self.0.serialize_unchecked(&mut writer)``, the *uncompressed* form
  of ``ProvingKey`` (groth16.rs:163-179), decoded with ``deserialize_unchecked`` (groth16.rs:142-160).
  Loading those bytes gives the five MSM base vectors the device library keeps resident.
* ``VerifyingContext`` serializes ``vk`` first (groth16.rs:200-212); the ``VerifyingKey`` part is
  handled here, the prepared (pairing) parts stay with arkworks on the host.

Encoding rules restated from ark-serialize / ark-ff / ark-ec 0.3.0 (crates.io dependencies that are
not vendored under /root/reference; **format recalled, not byte-compared with a real artefact** --
no Rust toolchain exists in this image):

* ``Fp``: canonical (non-Montgomery) integer, little-endian, ``ceil(MODULUS_BITS / 8)`` bytes
  (32 for the 254-bit fields, 48 for BLS12-381 Fq); flag bits live in the top bits of the last byte.
* ``Fp2`` (``QuadExtField``): ``c0`` then ``c1``; flags go on ``c1``.
* ``GroupAffine`` compressed: ``x`` with ``SWFlags`` -- bit 7 = "y is the larger of (y, -y)",
  bit 6 = infinity (then x = 0).  ``Fp`` orders by canonical integer, ``Fp2`` by ``c1`` then ``c0``.
* ``GroupAffine`` uncompressed: ``x``, then ``y`` carrying only the infinity flag; the point at
  infinity is written as (0, 1) + flag like ``GroupAffine::zero()``.
* ``Vec<T>``: u64 little-endian length, then the elements.
* ``Proof {a, b, c}``; ``VerifyingKey {alpha_g1, beta_g2, gamma_g2, delta_g2, gamma_abc_g1}``;
  ``ProvingKey {vk, beta_g1, delta_g1, a_query, b_g1_query, b_g2_query, h_query, l_query}``
  in declaration order (field names as destructured at groth16.rs:200-205).

Points cross this module in the C ABI's layout (``include/ozl.h``): little-endian u64 limbs in
Montgomery form, ``x || y`` (G2: ``x.c0 || x.c1 || y.c0 || y.c1``), infinity as a separate bit.
Host-only code: nothing here is on the timed path.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np

FLAG_Y_LARGER = 0x80     # SWFlags::PositiveY
FLAG_INFINITY = 0x40     # SWFlags::Infinity


class SerializationError(ValueError):
    """ark_serialize::SerializationError (InvalidData / UnexpectedFlags / NotEnoughSpace)."""


@dataclass(frozen=True)
class GroupSpec:
    name: str
    p: int          # base-field modulus
    degree: int     # 1 = Fq, 2 = Fq2 = Fq[u]/(u^2 + 1)
    b: tuple        # curve coefficient as `degree` canonical ints (y^2 = x^3 + b)

    @property
    def fq_bytes(self) -> int:
        return (self.p.bit_length() + 7) // 8

    @property
    def limbs(self) -> int:      # u64 limbs per Fq element
        return (self.p.bit_length() + 63) // 64

    @property
    def compressed_size(self) -> int:
        return self.degree * self.fq_bytes

    @property
    def uncompressed_size(self) -> int:
        return 2 * self.degree * self.fq_bytes


_Q381 = 0x1A0111EA397FE69A4B1BA7B6434BACD764774B84F38512BF6730D2A0F6B0F6241EABFFFEB153FFFFB9FEFFFFFFFFAAAB
_Q254 = 21888242871839275222246405745257275088696311157297823662689037894645226208583

BLS12_381_G1 = GroupSpec("bls12_381_g1", _Q381, 1, (4,))
BLS12_381_G2 = GroupSpec("bls12_381_g2", _Q381, 2, (4, 4))
BN254_G1 = GroupSpec("bn254_g1", _Q254, 1, (3,))
BN254_G2 = GroupSpec("bn254_g2", _Q254, 2,
                     (19485874751759354771024239261021720505790618469301721065564631296452457478373,
                      266929791119991161246907387137283842545076965332900288569378510910307636690))
GROUPS = {g.name: g for g in (BLS12_381_G1, BLS12_381_G2, BN254_G1, BN254_G2)}
PAIRING_GROUPS = {"bn254": (BN254_G1, BN254_G2), "bls12_381": (BLS12_381_G1, BLS12_381_G2)}


# --------------------------------------------------------------------------------------------
# base-field helpers (canonical ints; Fq2 elements are (c0, c1))
# --------------------------------------------------------------------------------------------
def _fq_sqrt(a: int, p: int) -> Optional[int]:
    # both base fields have p = 3 (mod 4)
    r = pow(a, (p + 1) // 4, p)
    return r if (r * r - a) % p == 0 else None


def _fq2_mul(a, b, p):
    return ((a[0] * b[0] - a[1] * b[1]) % p, (a[0] * b[1] + a[1] * b[0]) % p)


def _fq2_sqrt(a, p) -> Optional[tuple]:
    """Square root in Fq[u]/(u^2+1) by the norm method; None for a non-residue."""
    a0, a1 = a[0] % p, a[1] % p
    if a1 == 0:
        r = _fq_sqrt(a0, p)
        if r is not None:
            return (r, 0)
        r = _fq_sqrt((-a0) % p, p)          # sqrt(a0) = sqrt(-a0) * u
        return None if r is None else (0, r)
    s = _fq_sqrt((a0 * a0 + a1 * a1) % p, p)
    if s is None:
        return None
    inv2 = (p + 1) // 2
    for t in (s, (-s) % p):
        x0 = _fq_sqrt(((a0 + t) * inv2) % p, p)
        if x0 is not None and x0 != 0:
            x1 = (a1 * pow(2 * x0, -1, p)) % p
            if _fq2_mul((x0, x1), (x0, x1), p) == (a0, a1):
                return (x0, x1)
    return None


def _neg(v: tuple, p: int) -> tuple:
    return tuple((-c) % p for c in v)


def _is_larger(y: tuple, p: int) -> bool:
    """ark's ``y > -y``: Fp compares canonical integers; QuadExtField compares c1 first, then c0."""
    ny = _neg(y, p)
    return tuple(reversed(y)) > tuple(reversed(ny))


def _rhs(g: GroupSpec, x: tuple) -> tuple:
    p = g.p
    if g.degree == 1:
        return ((x[0] * x[0] * x[0] + g.b[0]) % p,)
    x3 = _fq2_mul(_fq2_mul(x, x, p), x, p)
    return ((x3[0] + g.b[0]) % p, (x3[1] + g.b[1]) % p)


def is_on_curve(g: GroupSpec, pt: Optional[Tuple[tuple, tuple]]) -> bool:
    if pt is None:
        return True
    x, y = pt
    y2 = ((y[0] * y[0]) % g.p,) if g.degree == 1 else _fq2_mul(y, y, g.p)
    return y2 == _rhs(g, x)


# --------------------------------------------------------------------------------------------
# field elements and points <-> bytes (points here are canonical tuples or None for infinity)
# --------------------------------------------------------------------------------------------
def _fe_to_bytes(g: GroupSpec, v: tuple, flags: int = 0) -> bytes:
    out = bytearray()
    for c in v:
        out += int(c).to_bytes(g.fq_bytes, "little")
    out[-1] |= flags
    return bytes(out)


def _fe_from_bytes(g: GroupSpec, buf: bytes, off: int, flag_mask: int) -> Tuple[tuple, int]:
    nb = g.fq_bytes
    end = off + g.degree * nb
    if end > len(buf):
        raise SerializationError("not enough bytes")
    raw = bytearray(buf[off:end])
    flags = raw[-1] & flag_mask
    raw[-1] &= (~flag_mask) & 0xFF
    v = tuple(int.from_bytes(raw[i * nb:(i + 1) * nb], "little") for i in range(g.degree))
    if any(c >= g.p for c in v):
        raise SerializationError("field element not below the modulus")
    return v, flags


def point_to_bytes(g: GroupSpec, pt: Optional[Tuple[tuple, tuple]], compressed: bool = True) -> bytes:
    zero = (0,) * g.degree
    if compressed:
        if pt is None:
            return _fe_to_bytes(g, zero, FLAG_INFINITY)
        x, y = pt
        return _fe_to_bytes(g, x, FLAG_Y_LARGER if _is_larger(y, g.p) else 0)
    if pt is None:
        one = (1,) + (0,) * (g.degree - 1)
        return _fe_to_bytes(g, zero) + _fe_to_bytes(g, one, FLAG_INFINITY)
    return _fe_to_bytes(g, pt[0]) + _fe_to_bytes(g, pt[1])


def point_from_bytes(g: GroupSpec, buf: bytes, off: int = 0, compressed: bool = True,
                     check: bool = True) -> Tuple[Optional[Tuple[tuple, tuple]], int]:
    """Returns (point, new offset).  ``check`` = on-curve validation of uncompressed input
    (``deserialize_unchecked`` skips it; subgroup membership is never checked here)."""
    if compressed:
        x, flags = _fe_from_bytes(g, buf, off, FLAG_Y_LARGER | FLAG_INFINITY)
        off += g.compressed_size
        if flags & FLAG_INFINITY:
            if flags & FLAG_Y_LARGER:
                raise SerializationError("unexpected flags")
            return None, off
        rhs = _rhs(g, x)
        y0 = _fq_sqrt(rhs[0], g.p) if g.degree == 1 else _fq2_sqrt(rhs, g.p)
        if y0 is None:
            raise SerializationError("x is not on the curve")
        y = (y0,) if g.degree == 1 else y0
        if _is_larger(y, g.p) != bool(flags & FLAG_Y_LARGER):
            y = _neg(y, g.p)
        return (x, y), off
    x, _ = _fe_from_bytes(g, buf, off, 0)
    y, flags = _fe_from_bytes(g, buf, off + g.compressed_size, FLAG_Y_LARGER | FLAG_INFINITY)
    off += g.uncompressed_size
    if flags & FLAG_INFINITY:
        return None, off
    if flags & FLAG_Y_LARGER:
        raise SerializationError("unexpected flags")
    if check and not is_on_curve(g, (x, y)):
        raise SerializationError("point is not on the curve")
    return (x, y), off


# --------------------------------------------------------------------------------------------
# C-ABI layout (Montgomery u64 limbs) <-> canonical tuples
# --------------------------------------------------------------------------------------------
def limbs_to_point(g: GroupSpec, limbs: np.ndarray, infinity: bool = False) -> Optional[Tuple[tuple, tuple]]:
    """One affine point in ABI layout (2 * degree * limbs u64, Montgomery) -> canonical tuple."""
    if infinity:
        return None
    L = g.limbs
    raw = np.ascontiguousarray(limbs, dtype=np.uint64).reshape(-1).tobytes()
    rinv = pow(1 << (64 * L), -1, g.p)
    c = [(int.from_bytes(raw[8 * L * i:8 * L * (i + 1)], "little") * rinv) % g.p for i in range(2 * g.degree)]
    return tuple(c[:g.degree]), tuple(c[g.degree:])


def point_to_limbs(g: GroupSpec, pt: Optional[Tuple[tuple, tuple]]) -> np.ndarray:
    L = g.limbs
    out = np.zeros(2 * g.degree * L, dtype=np.uint64)
    if pt is None:
        return out
    R = 1 << (64 * L)
    raw = b"".join(((int(c) * R) % g.p).to_bytes(8 * L, "little") for c in (*pt[0], *pt[1]))
    return np.frombuffer(raw, dtype=np.uint64).copy()


def points_to_limbs(g: GroupSpec, pts: Sequence[Optional[Tuple[tuple, tuple]]]) -> Tuple[np.ndarray, np.ndarray]:
    """Canonical points -> ((n, 2*degree*limbs) uint64 Montgomery, packed infinity bitset) as
    ``ozl_msm_bases_upload`` takes them."""
    n = len(pts)
    L = g.limbs
    R = 1 << (64 * L)
    p = g.p
    nb = 8 * L
    buf = bytearray(n * 2 * g.degree * nb)
    inf = np.zeros(n, dtype=np.uint8)
    o = 0
    for i, pt in enumerate(pts):
        if pt is None:
            inf[i] = 1
            o += 2 * g.degree * nb
            continue
        for c in (*pt[0], *pt[1]):
            buf[o:o + nb] = ((c * R) % p).to_bytes(nb, "little")
            o += nb
    arr = np.frombuffer(bytes(buf), dtype=np.uint64).reshape(n, 2 * g.degree * L).copy()
    return arr, np.packbits(inf, bitorder="little")


def limbs_to_points(g: GroupSpec, arr: np.ndarray, inf_mask: Optional[np.ndarray] = None) -> List[Optional[Tuple[tuple, tuple]]]:
    arr = np.ascontiguousarray(arr, dtype=np.uint64)
    n = arr.shape[0]
    flags = np.zeros(n, dtype=np.uint8) if inf_mask is None else np.unpackbits(
        np.ascontiguousarray(inf_mask, dtype=np.uint8), bitorder="little")[:n]
    return [limbs_to_point(g, arr[i], bool(flags[i])) for i in range(n)]


# --------------------------------------------------------------------------------------------
# Vec<GroupAffine>
# --------------------------------------------------------------------------------------------
def _vec_to_bytes(g: GroupSpec, pts, compressed: bool) -> bytes:
    return len(pts).to_bytes(8, "little") + b"".join(point_to_bytes(g, pt, compressed) for pt in pts)


def _vec_from_bytes(g: GroupSpec, buf: bytes, off: int, compressed: bool, check: bool):
    if off + 8 > len(buf):
        raise SerializationError("not enough bytes")
    n = int.from_bytes(buf[off:off + 8], "little")
    off += 8
    size = g.compressed_size if compressed else g.uncompressed_size
    if off + n * size > len(buf):
        raise SerializationError("vector length exceeds the buffer")
    pts = []
    for _ in range(n):
        pt, off = point_from_bytes(g, buf, off, compressed, check)
        pts.append(pt)
    return pts, off


# --------------------------------------------------------------------------------------------
# Proof
# --------------------------------------------------------------------------------------------
def proof_as_bytes(pairing: str, a, b, c) -> bytes:
    """``proof_as_bytes`` (groth16.rs:98-107): compressed a (G1) || b (G2) || c (G1).
    Arguments are canonical points (or None); 128 bytes for BN254, 192 for BLS12-381."""
    g1, g2 = PAIRING_GROUPS[pairing]
    return point_to_bytes(g1, a) + point_to_bytes(g2, b) + point_to_bytes(g1, c)


def proof_from_bytes(pairing: str, buf: bytes):
    """``Proof::try_from(Vec<u8>)`` (groth16.rs:86-92): returns canonical (a, b, c)."""
    g1, g2 = PAIRING_GROUPS[pairing]
    a, off = point_from_bytes(g1, buf, 0)
    b, off = point_from_bytes(g2, buf, off)
    c, off = point_from_bytes(g1, buf, off)
    return a, b, c


def proof_limbs_as_bytes(pairing: str, a: np.ndarray, b: np.ndarray, c: np.ndarray) -> bytes:
    """Same, from the three affine outputs of ``ozl_groth16_prove`` (Montgomery limbs; an all-zero
    output is the library's encoding of the identity)."""
    g1, g2 = PAIRING_GROUPS[pairing]

    def conv(g, v):
        v = np.ascontiguousarray(v, dtype=np.uint64).reshape(-1)
        return limbs_to_point(g, v, infinity=not v.any())
    return proof_as_bytes(pairing, conv(g1, a), conv(g2, b), conv(g1, c))


# --------------------------------------------------------------------------------------------
# VerifyingKey / ProvingKey
# --------------------------------------------------------------------------------------------
@dataclass
class VerifyingKey:
    alpha_g1: object
    beta_g2: object
    gamma_g2: object
    delta_g2: object
    gamma_abc_g1: list = field(default_factory=list)


@dataclass
class ProvingKey:
    """``ark_groth16::ProvingKey`` as canonical points; ``*_query`` are the five MSM base vectors."""
    vk: VerifyingKey
    beta_g1: object
    delta_g1: object
    a_query: list
    b_g1_query: list
    b_g2_query: list
    h_query: list
    l_query: list


def vk_to_bytes(pairing: str, vk: VerifyingKey, compressed: bool = True) -> bytes:
    g1, g2 = PAIRING_GROUPS[pairing]
    return (point_to_bytes(g1, vk.alpha_g1, compressed) + point_to_bytes(g2, vk.beta_g2, compressed)
            + point_to_bytes(g2, vk.gamma_g2, compressed) + point_to_bytes(g2, vk.delta_g2, compressed)
            + _vec_to_bytes(g1, vk.gamma_abc_g1, compressed))


def vk_from_bytes(pairing: str, buf: bytes, off: int = 0, compressed: bool = True, check: bool = True):
    g1, g2 = PAIRING_GROUPS[pairing]
    alpha, off = point_from_bytes(g1, buf, off, compressed, check)
    beta, off = point_from_bytes(g2, buf, off, compressed, check)
    gamma, off = point_from_bytes(g2, buf, off, compressed, check)
    delta, off = point_from_bytes(g2, buf, off, compressed, check)
    ic, off = _vec_from_bytes(g1, buf, off, compressed, check)
    return VerifyingKey(alpha, beta, gamma, delta, ic), off


def proving_key_to_bytes(pairing: str, pk: ProvingKey, compressed: bool = False) -> bytes:
    """``ProvingContext::encode`` = ``serialize_unchecked`` (groth16.rs:163-179) when compressed=False."""
    g1, g2 = PAIRING_GROUPS[pairing]
    return (vk_to_bytes(pairing, pk.vk, compressed)
            + point_to_bytes(g1, pk.beta_g1, compressed) + point_to_bytes(g1, pk.delta_g1, compressed)
            + _vec_to_bytes(g1, pk.a_query, compressed) + _vec_to_bytes(g1, pk.b_g1_query, compressed)
            + _vec_to_bytes(g2, pk.b_g2_query, compressed) + _vec_to_bytes(g1, pk.h_query, compressed)
            + _vec_to_bytes(g1, pk.l_query, compressed))


def proving_key_from_bytes(pairing: str, buf: bytes, compressed: bool = False, check: bool = False) -> ProvingKey:
    """``ProvingContext::decode`` = ``deserialize_unchecked`` (groth16.rs:142-160): uncompressed and
    without curve checks by default, like the reference.  Trailing bytes are an error
    (``ArkReader::finish``, serialize.rs:33-169)."""
    g1, g2 = PAIRING_GROUPS[pairing]
    vk, off = vk_from_bytes(pairing, buf, 0, compressed, check)
    beta_g1, off = point_from_bytes(g1, buf, off, compressed, check)
    delta_g1, off = point_from_bytes(g1, buf, off, compressed, check)
    a_q, off = _vec_from_bytes(g1, buf, off, compressed, check)
    b1_q, off = _vec_from_bytes(g1, buf, off, compressed, check)
    b2_q, off = _vec_from_bytes(g2, buf, off, compressed, check)
    h_q, off = _vec_from_bytes(g1, buf, off, compressed, check)
    l_q, off = _vec_from_bytes(g1, buf, off, compressed, check)
    if off != len(buf):
        raise SerializationError("trailing bytes after the proving key")
    return ProvingKey(vk, beta_g1, delta_g1, a_q, b1_q, b2_q, h_q, l_q)
