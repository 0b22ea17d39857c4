"""Wire formats: ark-serialize 0.3.0 ``CanonicalSerialize`` as used by
``proof_as_bytes`` / ``ProvingContext::{encode, decode}``
(/root/reference/plugins/arkworks/src/groth16.rs:98-107, 142-179).  CPU only; the oracle's curve
arithmetic supplies points and checks what the decoder reconstructs."""
import random

import numpy as np
import pytest

from openzl_b200 import serialize as ser
from oracle import curves

PAIRS = [("bn254_g1", ser.BN254_G1), ("bn254_g2", ser.BN254_G2),
         ("bls12_381_g1", ser.BLS12_381_G1), ("bls12_381_g2", ser.BLS12_381_G2)]


def _to_ser(g, P):
    """oracle point -> serialize-module point (coordinates as tuples of Fq ints)."""
    if P is None:
        return None
    return ((P[0],), (P[1],)) if g.degree == 1 else (tuple(P[0]), tuple(P[1]))


def _points(name, g, k, seed):
    C = curves.CURVES[name]
    rnd = random.Random(seed)
    return [_to_ser(g, C.mul_affine(C.gen, rnd.randrange(1, C.fr.p))) for _ in range(k)]


def test_sizes_match_ark():
    assert ser.BN254_G1.compressed_size == 32 and ser.BN254_G2.compressed_size == 64
    assert ser.BLS12_381_G1.compressed_size == 48 and ser.BLS12_381_G2.compressed_size == 96
    assert ser.BN254_G1.uncompressed_size == 64 and ser.BLS12_381_G2.uncompressed_size == 192


def test_bn254_generator_known_bytes():
    # (1, 2): -y = q - 2 is the larger root, so no sign flag; x = 1 little-endian
    b = ser.point_to_bytes(ser.BN254_G1, ((1,), (2,)))
    assert b == bytes([1] + [0] * 31)
    neg = ser.point_to_bytes(ser.BN254_G1, ((1,), (ser.BN254_G1.p - 2,)))
    assert neg == bytes([1] + [0] * 30 + [0x80])
    inf = ser.point_to_bytes(ser.BN254_G1, None)
    assert inf == bytes([0] * 31 + [0x40])
    unc = ser.point_to_bytes(ser.BN254_G1, None, compressed=False)
    assert unc == bytes([0] * 32 + [1] + [0] * 30 + [0x40])        # GroupAffine::zero() = (0, 1, inf)


@pytest.mark.parametrize("name,g", PAIRS)
def test_point_round_trip(name, g):
    C = curves.CURVES[name]
    pts = _points(name, g, 6, 11) + [None]
    pts.append(_to_ser(g, C.neg(C.gen)))
    for compressed in (True, False):
        for pt in pts:
            b = ser.point_to_bytes(g, pt, compressed)
            assert len(b) == (g.compressed_size if compressed else g.uncompressed_size)
            back, off = ser.point_from_bytes(g, b, 0, compressed)
            assert off == len(b) and back == pt
            assert ser.is_on_curve(g, back)


@pytest.mark.parametrize("name,g", PAIRS)
def test_sign_flag_selects_the_larger_root(name, g):
    C = curves.CURVES[name]
    P = _to_ser(g, C.gen)
    N = _to_ser(g, C.neg(C.gen))
    bp, bn = ser.point_to_bytes(g, P), ser.point_to_bytes(g, N)
    assert bp[:-1] == bn[:-1] and (bp[-1] ^ bn[-1]) == ser.FLAG_Y_LARGER
    larger = P if (bp[-1] & ser.FLAG_Y_LARGER) else N
    other = N if larger is P else P
    key = (lambda y: tuple(reversed(y)))          # Fp2 orders by c1 first
    assert key(larger[1]) > key(other[1])


@pytest.mark.parametrize("name,g", PAIRS)
def test_rejects_bad_input(name, g):
    with pytest.raises(ser.SerializationError):
        ser.point_from_bytes(g, bytes(g.compressed_size - 1))
    # x with no square root on the curve
    C = curves.CURVES[name]
    x = 1
    while True:
        cand = ((x,), ) if g.degree == 1 else ((x, 0),)
        rhs = ser._rhs(g, cand[0])
        root = ser._fq_sqrt(rhs[0], g.p) if g.degree == 1 else ser._fq2_sqrt(rhs, g.p)
        if root is None:
            break
        x += 1
    with pytest.raises(ser.SerializationError):
        ser.point_from_bytes(g, ser._fe_to_bytes(g, cand[0]))
    # coordinate >= modulus
    raw = bytearray((g.p).to_bytes(g.fq_bytes, "little") * g.degree)
    with pytest.raises(ser.SerializationError):
        ser.point_from_bytes(g, bytes(raw))
    # infinity + sign flag together
    bad = bytearray(g.compressed_size)
    bad[-1] = 0xC0
    with pytest.raises(ser.SerializationError):
        ser.point_from_bytes(g, bytes(bad))
    # uncompressed point off the curve is caught only when checking
    P = _to_ser(g, C.gen)
    off_curve = (P[0], ser._neg(P[0], g.p))
    b = ser.point_to_bytes(g, off_curve, compressed=False)
    assert ser.point_from_bytes(g, b, 0, False, check=False)[0] == off_curve
    with pytest.raises(ser.SerializationError):
        ser.point_from_bytes(g, b, 0, False, check=True)


def test_fq2_sqrt():
    rnd = random.Random(5)
    for p in (ser.BN254_G2.p, ser.BLS12_381_G2.p):
        found = 0
        for _ in range(40):
            a = (rnd.randrange(p), rnd.randrange(p))
            sq = ser._fq2_mul(a, a, p)
            r = ser._fq2_sqrt(sq, p)
            assert r is not None and ser._fq2_mul(r, r, p) == sq
            r2 = ser._fq2_sqrt(a, p)
            if r2 is not None:
                found += 1
                assert ser._fq2_mul(r2, r2, p) == a
        assert 5 < found < 35                      # about half the elements are squares
        for a0 in (4, p - 4, 3, p - 3):            # c1 = 0 branch, residues and non-residues of Fq
            r = ser._fq2_sqrt((a0, 0), p)
            assert r is not None and ser._fq2_mul(r, r, p) == (a0, 0)   # every Fq element is a square in Fq2


@pytest.mark.parametrize("pairing", ["bn254", "bls12_381"])
def test_proof_bytes(pairing):
    g1, g2 = ser.PAIRING_GROUPS[pairing]
    a, c = _points(g1.name, g1, 2, 3)
    b = _points(g2.name, g2, 1, 4)[0]
    raw = ser.proof_as_bytes(pairing, a, b, c)
    assert len(raw) == (128 if pairing == "bn254" else 192)
    assert ser.proof_from_bytes(pairing, raw) == (a, b, c)
    # from the prover's output layout (Montgomery limbs)
    la, lb, lc = (ser.point_to_limbs(g, p) for g, p in ((g1, a), (g2, b), (g1, c)))
    assert ser.proof_limbs_as_bytes(pairing, la, lb, lc) == raw
    C1 = curves.CURVES[g1.name]
    assert list(la) == C1.affine_to_mont_limbs((a[0][0], a[1][0]))       # ABI layout agrees with the oracle's
    # identity output (all-zero limbs) serializes as the infinity encoding
    z = np.zeros_like(la)
    assert ser.proof_limbs_as_bytes(pairing, z, lb, lc)[:g1.compressed_size] == ser.point_to_bytes(g1, None)


@pytest.mark.parametrize("pairing", ["bn254", "bls12_381"])
def test_proving_key_round_trip(pairing):
    g1, g2 = ser.PAIRING_GROUPS[pairing]
    p1 = _points(g1.name, g1, 16, 7)
    p2 = _points(g2.name, g2, 8, 8)
    vk = ser.VerifyingKey(p1[0], p2[0], p2[1], p2[2], [p1[1], p1[2]])
    pk = ser.ProvingKey(vk, p1[3], p1[4], a_query=p1[5:9], b_g1_query=[p1[9], None, p1[10], None],
                        b_g2_query=[p2[3], None, p2[4], None], h_query=p1[11:14], l_query=p1[14:16])
    raw = ser.proving_key_to_bytes(pairing, pk)                    # ProvingContext::encode (uncompressed)
    s1, s2 = g1.uncompressed_size, g2.uncompressed_size
    assert len(raw) == (s1 + 3 * s2 + 8 + 2 * s1) + 2 * s1 + (8 + 4 * s1) * 2 + (8 + 4 * s2) + (8 + 3 * s1) + (8 + 2 * s1)
    back = ser.proving_key_from_bytes(pairing, raw)
    assert back == pk
    with pytest.raises(ser.SerializationError):
        ser.proving_key_from_bytes(pairing, raw + b"\0")
    with pytest.raises(ser.SerializationError):
        ser.proving_key_from_bytes(pairing, raw[:-1])
    comp = ser.proving_key_to_bytes(pairing, pk, compressed=True)
    assert len(comp) < len(raw) and ser.proving_key_from_bytes(pairing, comp, compressed=True) == pk
    # bases in the ABI layout, with the infinity bitset ozl_msm_bases_upload takes
    arr, mask = ser.points_to_limbs(g2, back.b_g2_query)
    assert arr.shape == (4, 4 * g2.limbs) and list(np.unpackbits(mask, bitorder="little")[:4]) == [0, 1, 0, 1]
    assert ser.limbs_to_points(g2, arr, mask) == back.b_g2_query
    vkb = ser.vk_to_bytes(pairing, vk)
    assert ser.vk_from_bytes(pairing, vkb)[0] == vk


def test_random_byte_strings_never_crash_the_decoder():
    """``deserialize`` on arbitrary bytes either returns a point on the curve or raises
    SerializationError (ark: Err(SerializationError)); nothing else."""
    rnd = random.Random(2026)
    for _, g in PAIRS:
        ok = 0
        for _ in range(60):
            raw = bytes(rnd.getrandbits(8) for _ in range(g.compressed_size))
            try:
                pt, off = ser.point_from_bytes(g, raw)
            except ser.SerializationError:
                continue
            ok += 1
            assert off == g.compressed_size and ser.is_on_curve(g, pt)
            if pt is not None:
                assert ser.point_to_bytes(g, pt) == raw          # canonical: re-encoding reproduces the bytes
        assert ok < 60


def test_limb_layout_matches_the_abi_for_g2():
    """x.c0 || x.c1 || y.c0 || y.c1, each a little-endian Montgomery residue (include/ozl.h)."""
    g = ser.BN254_G2
    C = curves.CURVES["bn254_g2"]
    P = _to_ser(g, C.gen)
    limbs = ser.point_to_limbs(g, P)
    assert list(limbs) == C.affine_to_mont_limbs(C.gen)
    assert ser.limbs_to_point(g, limbs) == P
