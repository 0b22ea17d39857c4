"""Host-side pairing and Groth16 verification: the product's mirror of the plugin's ``Pairing`` /
``PairingEngineExt`` (/root/reference/plugins/arkworks/src/pairing.rs:9-90) and of
``ProofSystem::verify`` (/root/reference/plugins/arkworks/src/groth16.rs:459-466, which calls
``ark_groth16::verify_with_processed_vk``).

Verification stays on the HOST in the reference (three Miller loops and one final exponentiation,
milliseconds) and SURVEY.md row a-6 keeps it there: nothing here touches the GPU.  It exists so that a
user of the device prover can check what it produced without arkworks.

Algorithm: the ate pairing  a(P, Q) = f_{T,Q}(P)^((p^12 - 1)/r)  with T = t - 1 (trace of Frobenius minus
one: 6x^2 for BN254, the curve parameter x for BLS12-381, where it IS the optimal ate loop).  Q lives on
the sextic twist E'(Fq2) and the line functions are evaluated through the untwisting map into
Fq12 = Fq2[w]/(w^6 - xi):

    D-type twist (BN254,     b' = b / xi):  psi(x, y) = (x w^2, y w^3)
        line = yP - (lambda xP) w + (lambda xT - yT) w^3
    M-type twist (BLS12-381, b' = b * xi):  psi(x, y) = (x / w^2, y / w^3)
        line * w^3 = (lambda xT - yT) - (lambda xP) w^2 + yP w^3

Vertical lines, the sign of T and the factor w^3 all lie in proper subfields (or amount to a fixed power
of the pairing) and vanish in the final exponentiation / do not matter for an equation between products
of pairings.  The value differs from arkworks' optimal-ate value by a fixed exponent, like any two
pairings on the same groups; verification decisions are identical.  The final exponentiation is a plain
square-and-multiply over the 2.8-kbit exponent: ~0.3 s in CPython, once per verification.

This file shares no code with the test suite's checker (a reduced Tate pairing over Fq[w]/(m(w)));
tests/test_pairing_host.py checks the two against each other and against bilinearity.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np

Fq2 = Tuple[int, int]


@dataclass(frozen=True)
class Engine:
    name: str
    p: int            # base field
    r: int            # group order
    b1: int           # E : y^2 = x^3 + b1 over Fq
    xi: Fq2           # Fq12 = Fq2[w]/(w^6 - xi)
    twist: str        # "D" or "M"
    loop: int         # |t - 1|
    g1: Tuple[int, int]
    g2: Tuple[Fq2, Fq2]


_BN_P = 21888242871839275222246405745257275088696311157297823662689037894645226208583
_BN_R = 21888242871839275222246405745257275088548364400416034343698204186575808495617
_BLS_P = 0x1A0111EA397FE69A4B1BA7B6434BACD764774B84F38512BF6730D2A0F6B0F6241EABFFFEB153FFFFB9FEFFFFFFFFAAAB
_BLS_R = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001

BN254 = Engine(
    "bn254", _BN_P, _BN_R, 3, (9, 1), "D", _BN_P - _BN_R, (1, 2),
    ((10857046999023057135944570762232829481370756359578518086990519993285655852781,
      11559732032986387107991004021392285783925812861821192530917403151452391805634),
     (8495653923123431417604973247489272438418190587263600148770280649306958101930,
      4082367875863433681332203403145435568316851327593401208105741076214120093531)))
BLS12_381 = Engine(
    "bls12_381", _BLS_P, _BLS_R, 4, (1, 1), "M", 0xD201000000010000,
    (3685416753713387016781088315183077757961620795782546409894578378688607592378376318836054947676345821548104185464507,
     1339506544944476473020471379941921221584933875938349620426543736416511423956333506472724655353366534992391756441569),
    ((352701069587466618187139116011060144890029952792775240219908644239793785735715026873347600343865175952761926303160,
      3059144344244213709971259814753781636986470325476647558659373206291635324768958432433509563104347017837885763365758),
     (1985150602287291935568054521177171638300868978215655730859378665066344726373823718423869104263333984641494340347905,
      927553665492332455747201965776037880757740193453592970025027978793976877002675564980949289727957565575433344219582)))
ENGINES = {"bn254": BN254, "bls12_381": BLS12_381}


# ---------------------------------------------------------------------------------------------
# Fq2 = Fq[u]/(u^2 + 1)
# ---------------------------------------------------------------------------------------------
def _f2_add(a, b, p):
    return ((a[0] + b[0]) % p, (a[1] + b[1]) % p)


def _f2_sub(a, b, p):
    return ((a[0] - b[0]) % p, (a[1] - b[1]) % p)


def _f2_mul(a, b, p):
    t0, t1 = a[0] * b[0], a[1] * b[1]
    return ((t0 - t1) % p, ((a[0] + a[1]) * (b[0] + b[1]) - t0 - t1) % p)


def _f2_scale(a, k, p):
    return (a[0] * k % p, a[1] * k % p)


def _f2_inv(a, p):
    n = pow(a[0] * a[0] + a[1] * a[1], -1, p)
    return (a[0] * n % p, -a[1] * n % p)


def _f2_neg(a, p):
    return (-a[0] % p, -a[1] % p)


# ---------------------------------------------------------------------------------------------
# Fq12 = Fq2[w]/(w^6 - xi): six Fq2 coefficients, low degree first
# ---------------------------------------------------------------------------------------------
_ZERO2 = (0, 0)


def _f12_one():
    return [(1, 0)] + [_ZERO2] * 5


def _f12_mul(a, b, E: Engine):
    p = E.p
    t = [[0, 0] for _ in range(11)]
    for i, ai in enumerate(a):
        if ai == _ZERO2:
            continue
        for j, bj in enumerate(b):
            if bj == _ZERO2:
                continue
            m = _f2_mul(ai, bj, p)
            t[i + j][0] += m[0]
            t[i + j][1] += m[1]
    out = []
    for k in range(6):
        c0, c1 = t[k]
        if k + 6 < 11:
            hi = _f2_mul((t[k + 6][0] % p, t[k + 6][1] % p), E.xi, p)     # w^(k+6) = xi w^k
            c0 += hi[0]
            c1 += hi[1]
        out.append((c0 % p, c1 % p))
    return out


def _f12_pow(a, e: int, E: Engine):
    res = _f12_one()
    for bit in bin(e)[2:]:
        res = _f12_mul(res, res, E)
        if bit == "1":
            res = _f12_mul(res, a, E)
    return res


# ---------------------------------------------------------------------------------------------
# affine curve arithmetic (host; a handful of operations per verification)
# ---------------------------------------------------------------------------------------------
def _g1_add(P, Q, p):
    if P is None:
        return Q
    if Q is None:
        return P
    if P[0] == Q[0]:
        if (P[1] + Q[1]) % p == 0:
            return None
        lam = 3 * P[0] * P[0] * pow(2 * P[1], -1, p) % p
    else:
        lam = (Q[1] - P[1]) * pow(Q[0] - P[0], -1, p) % p
    x = (lam * lam - P[0] - Q[0]) % p
    return (x, (lam * (P[0] - x) - P[1]) % p)


def g1_mul(E: Engine, P, k: int):
    acc = None
    k %= E.r
    for bit in bin(k)[2:] if k else "":
        acc = _g1_add(acc, acc, E.p)
        if bit == "1":
            acc = _g1_add(acc, P, E.p)
    return acc


def g1_neg(E: Engine, P):
    return None if P is None else (P[0], -P[1] % E.p)


def _g2_add(P, Q, p):
    if P is None:
        return Q
    if Q is None:
        return P
    if P[0] == Q[0]:
        if _f2_add(P[1], Q[1], p) == _ZERO2:
            return None
        lam = _f2_mul(_f2_scale(_f2_mul(P[0], P[0], p), 3, p), _f2_inv(_f2_scale(P[1], 2, p), p), p)
    else:
        lam = _f2_mul(_f2_sub(Q[1], P[1], p), _f2_inv(_f2_sub(Q[0], P[0], p), p), p)
    x = _f2_sub(_f2_sub(_f2_mul(lam, lam, p), P[0], p), Q[0], p)
    return (x, _f2_sub(_f2_mul(lam, _f2_sub(P[0], x, p), p), P[1], p))


def g2_mul(E: Engine, Q, k: int):
    acc = None
    k %= E.r
    for bit in bin(k)[2:] if k else "":
        acc = _g2_add(acc, acc, E.p)
        if bit == "1":
            acc = _g2_add(acc, Q, E.p)
    return acc


def g1_on_curve(E: Engine, P) -> bool:
    return P is None or (P[1] * P[1] - P[0] ** 3 - E.b1) % E.p == 0


def _b2(E: Engine) -> Fq2:
    b = (E.b1, 0)
    return _f2_mul(b, _f2_inv(E.xi, E.p), E.p) if E.twist == "D" else _f2_mul(b, E.xi, E.p)


def g2_on_curve(E: Engine, Q) -> bool:
    if Q is None:
        return True
    p = E.p
    lhs = _f2_mul(Q[1], Q[1], p)
    rhs = _f2_add(_f2_mul(_f2_mul(Q[0], Q[0], p), Q[0], p), _b2(E), p)
    return lhs == rhs


# ---------------------------------------------------------------------------------------------
# Miller loop and pairing
# ---------------------------------------------------------------------------------------------
def _line(E: Engine, lam: Fq2, T, P):
    """The line of slope `lam` through the twist point T, evaluated at P in G1 (sparse Fq12 element)."""
    p = E.p
    c_const = _f2_sub(_f2_mul(lam, T[0], p), T[1], p)          # lambda xT - yT
    c_x = _f2_neg(_f2_scale(lam, P[0], p), p)                    # -lambda xP
    c_y = (P[1] % p, 0)                                          # yP
    out = [_ZERO2] * 6
    if E.twist == "D":
        out[0], out[1], out[3] = c_y, c_x, c_const
    else:
        out[0], out[2], out[3] = c_const, c_x, c_y
    return out


def miller_loop(E: Engine, P, Q) -> Optional[list]:
    """f_{|t-1|, Q}(P) without vertical lines.  P in E(Fq), Q in E'(Fq2), both finite."""
    p = E.p
    f = _f12_one()
    T = Q
    for bit in bin(E.loop)[3:]:
        if _f2_scale(T[1], 2, p) == _ZERO2:
            return None                                         # a 2-torsion point: not a G2 element
        lam = _f2_mul(_f2_scale(_f2_mul(T[0], T[0], p), 3, p), _f2_inv(_f2_scale(T[1], 2, p), p), p)
        f = _f12_mul(_f12_mul(f, f, E), _line(E, lam, T, P), E)
        T = _g2_add(T, T, p)
        if bit == "1":
            if T is None or T[0] == Q[0]:
                return None                                     # only for points outside the order-r subgroup
            lam = _f2_mul(_f2_sub(Q[1], T[1], p), _f2_inv(_f2_sub(Q[0], T[0], p), p), p)
            f = _f12_mul(f, _line(E, lam, T, P), E)
            T = _g2_add(T, Q, p)
        if T is None:
            return None
    return f


def final_exponentiation(E: Engine, f):
    return _f12_pow(f, (E.p ** 12 - 1) // E.r, E)


def pairing(E: Engine, P, Q):
    """a(P, Q) in mu_r (as an Fq12 element); the identity of either group maps to 1."""
    if P is None or Q is None:
        return _f12_one()
    f = miller_loop(E, P, Q)
    if f is None:
        raise ValueError("G2 input is not in the order-r subgroup")
    return final_exponentiation(E, f)


def product_of_pairings_is_one(E: Engine, pairs: Sequence[Tuple[object, object]]) -> bool:
    """``PairingEngine::product_of_pairings`` (pairing.rs:47-52) compared with one: the Miller-loop
    values are multiplied and exponentiated ONCE."""
    acc = _f12_one()
    for P, Q in pairs:
        if P is None or Q is None:
            continue
        f = miller_loop(E, P, Q)
        if f is None:
            return False
        acc = _f12_mul(acc, f, E)
    return final_exponentiation(E, acc) == _f12_one()


# ``PairingEngineExt`` of the plugin (/root/reference/plugins/arkworks/src/pairing.rs:46-90): same names, same meaning.
# A "pair" is (G1 point, G2 point) in the affine tuples used throughout this module (None = identity).
def eval(E: Engine, pair):   # noqa: A001  (the reference's name)
    """``PairingEngineExt::eval``: the pairing function on one pair (pairing.rs:49-52)."""
    return pairing(E, pair[0], pair[1])


def has_same(E: Engine, lhs, rhs) -> bool:
    """``PairingEngineExt::has_same``: both pairs evaluate to the same target-group element (pairing.rs:55-58)."""
    return eval(E, lhs) == eval(E, rhs)


def same(E: Engine, lhs, rhs):
    """``PairingEngineExt::same``: ``Some((lhs, rhs))`` when the pairs evaluate to the same element, i.e. when there is
    an r with r * lhs[0] == rhs[0] and lhs[1] == r * rhs[1]; ``None`` otherwise (pairing.rs:60-74)."""
    return (lhs, rhs) if has_same(E, lhs, rhs) else None


def same_ratio(E: Engine, lhs, rhs) -> bool:
    """``PairingEngineExt::same_ratio``: the ratio of the G1 pair ``lhs = (L1, R1)`` equals the ratio of the G2 pair
    ``rhs = (L2, R2)``: e(L1, R2) == e(R1, L2) (pairing.rs:76-88)."""
    return has_same(E, (lhs[0], rhs[1]), (lhs[1], rhs[0]))


# ---------------------------------------------------------------------------------------------
# C-ABI layouts -> canonical integers
# ---------------------------------------------------------------------------------------------
def _from_mont(limbs: np.ndarray, p: int, nlimbs: int) -> List[int]:
    raw = np.ascontiguousarray(limbs, dtype=np.uint64).tobytes()
    rinv = pow(1 << (64 * nlimbs), -1, p)
    step = 8 * nlimbs
    return [int.from_bytes(raw[i:i + step], "little") * rinv % p for i in range(0, len(raw), step)]


def g1_from_limbs(E: Engine, limbs) -> Optional[Tuple[int, int]]:
    """x || y Montgomery limbs (ozl.h layout) -> affine point; all-zero = the point at infinity."""
    a = np.asarray(limbs, dtype=np.uint64).reshape(-1)
    if not a.any():
        return None
    x, y = _from_mont(a, E.p, len(a) // 2)
    return (x, y)


def g2_from_limbs(E: Engine, limbs) -> Optional[Tuple[Fq2, Fq2]]:
    a = np.asarray(limbs, dtype=np.uint64).reshape(-1)
    if not a.any():
        return None
    x0, x1, y0, y1 = _from_mont(a, E.p, len(a) // 4)
    return ((x0, x1), (y0, y1))


def groth16_verify(pairing_name: str, alpha_g1, beta_g2, gamma_g2, delta_g2, gamma_abc_g1, public_inputs: Sequence[int],
                   proof_a, proof_b, proof_c) -> bool:
    """e(A, B) = e(alpha, beta) e(sum_j x_j gamma_abc_j, gamma) e(C, delta) with x_0 = 1, every point given
    as affine Montgomery limbs (the layout `ozl_groth16_prove` writes).  Malformed input -- wrong number of
    public inputs, a point off its curve -- is a rejected proof, not an exception, like the reference's
    ``Result<bool, Error>`` collapsed to a decision."""
    E = ENGINES[pairing_name]
    abc = [g1_from_limbs(E, row) for row in np.asarray(gamma_abc_g1, dtype=np.uint64)]
    if len(public_inputs) + 1 != len(abc):
        return False
    A, B, C = g1_from_limbs(E, proof_a), g2_from_limbs(E, proof_b), g1_from_limbs(E, proof_c)
    alpha, beta = g1_from_limbs(E, alpha_g1), g2_from_limbs(E, beta_g2)
    gamma, delta = g2_from_limbs(E, gamma_g2), g2_from_limbs(E, delta_g2)
    if not (g1_on_curve(E, A) and g2_on_curve(E, B) and g1_on_curve(E, C)):
        return False
    ic = abc[0]
    for x, pt in zip(public_inputs, abc[1:]):
        ic = _g1_add(ic, g1_mul(E, pt, int(x)), E.p)
    return product_of_pairings_is_one(E, [(A, B), (g1_neg(E, alpha), beta), (g1_neg(E, ic), gamma), (g1_neg(E, C), delta)])
