"""Pairings for the verification equation (oracle; test infrastructure only).

The reference verifies with ``E::product_of_pairings`` / ``verify_with_processed_vk``
(``/root/reference/plugins/arkworks/src/pairing.rs:47-90``, ``groth16.rs:460-466``): three Miller loops
and one final exponentiation on the CPU.  That stays on the host in the product (SURVEY section 8 row
a-6); this file is a slow, transparent big-int restatement used ONLY by the tests to check that proofs
satisfy  e(A, B) = e(alpha, beta) e(sum x_i gamma_abc_i, gamma) e(C, delta)  with real pairings instead
of the known-trapdoor shortcut.

It is the reduced Tate pairing  e(P, Q) = f_{r,P}(Q)^((p^12 - 1)/r)  with P in G1 = E(Fq)[r] and
Q in G2 given on the sextic twist and mapped into E(Fq12).  It is NOT arkworks' optimal-ate pairing
value (a fixed power of it), which does not matter for an equation between products of pairings:
both are non-degenerate bilinear maps G1 x G2 -> mu_r.  Fq12 is represented as Fq[w]/(m(w)) with
w^6 = xi, so Fq2 = Fq[u]/(u^2 + 1) embeds through u = w^6 - xi_0:
    BN254      xi = 9 + u   m(w) = w^12 - 18 w^6 + 82    D-type twist  (x', y') -> (x' w^2, y' w^3)
    BLS12-381  xi = 1 + u   m(w) = w^12 -  2 w^6 +  2    M-type twist  (x', y') -> (x' / w^2, y' / w^3)
Vertical lines are skipped: x_Q lies in the subfield Fq6, so they die in the final exponentiation.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

from .curves import BLS12_381_G1, BLS12_381_G2, BN254_G1, BN254_G2, Curve


@dataclass(frozen=True)
class PairingParams:
    name: str
    g1: Curve
    g2: Curve
    xi0: int          # xi = xi0 + u
    m6: int           # m(w) = w^12 + m6 w^6 + m0
    m0: int
    twist: str        # "D" or "M"


BN254 = PairingParams("bn254", BN254_G1, BN254_G2, 9, -18, 82, "D")
BLS12_381 = PairingParams("bls12_381", BLS12_381_G1, BLS12_381_G2, 1, -2, 2, "M")
PAIRINGS = {"bn254": BN254, "bls12_381": BLS12_381}


class Fq12:
    """Arithmetic in Fq[w]/(w^12 + m6 w^6 + m0); elements are lists of 12 ints (low degree first)."""

    def __init__(self, P: PairingParams):
        self.P = P
        self.p = P.g1.base.p
        self.one = [1] + [0] * 11

    def mul(self, a: Sequence[int], b: Sequence[int]) -> List[int]:
        p = self.p
        t = [0] * 23
        for i, ai in enumerate(a):
            if ai:
                for j, bj in enumerate(b):
                    if bj:
                        t[i + j] += ai * bj
        m6, m0 = self.P.m6, self.P.m0
        for k in range(22, 11, -1):          # w^k = -(m6 w^(k-6) + m0 w^(k-12))
            c = t[k]
            if c:
                t[k - 6] -= m6 * c
                t[k - 12] -= m0 * c
        return [v % p for v in t[:12]]

    def sqr(self, a):
        return self.mul(a, a)

    def pow(self, a, e: int):
        res = self.one
        for bit in bin(e)[2:]:
            res = self.sqr(res)
            if bit == "1":
                res = self.mul(res, a)
        return res

    def from_fq2(self, c: Tuple[int, int], shift: int = 0) -> List[int]:
        """(a + b u) * w^shift with u = w^6 - xi0, for shift < 6."""
        out = [0] * 12
        out[shift] = (c[0] - self.P.xi0 * c[1]) % self.p
        out[shift + 6] = c[1] % self.p
        return out

    def w_inverse_power(self, k: int) -> List[int]:
        """w^(-k): w^12 = -(m6 w^6 + m0)  =>  w^-1 = -(w^11 + m6 w^5) / m0."""
        p = self.p
        inv_m0 = pow(self.P.m0 % p, -1, p)
        winv = [0] * 12
        winv[11] = (-inv_m0) % p
        winv[5] = (-self.P.m6 * inv_m0) % p
        assert self.mul(winv, [0, 1] + [0] * 10) == self.one
        return self.pow(winv, k)


def untwist(P: PairingParams, F: Fq12, Q) -> Tuple[List[int], List[int]]:
    """G2 point on the twist (Fq2 coordinates) -> point of E(Fq12)."""
    x, y = Q
    if P.twist == "D":
        return F.from_fq2(x, 2), F.from_fq2(y, 3)
    return F.mul(F.from_fq2(x), F.w_inverse_power(2)), F.mul(F.from_fq2(y), F.w_inverse_power(3))


def miller_loop(P: PairingParams, F: Fq12, P1, Q2) -> List[int]:
    """f_{r,P1}(psi(Q2)) without vertical lines; ``None`` operands give 1."""
    if P1 is None or Q2 is None:
        return F.one
    p = F.p
    r = P.g1.fr.p
    xq, yq = untwist(P, F, Q2)
    px, py = P1

    def line(tx, ty, lam):
        # (y_Q - ty) - lam (x_Q - tx), lam in Fq
        out = [(yv - lam * xv) % p for xv, yv in zip(xq, yq)]
        out[0] = (out[0] - ty + lam * tx) % p
        return out

    f = F.one
    tx, ty = px, py
    for bit in bin(r)[3:]:
        lam = (3 * tx * tx) * pow(2 * ty, -1, p) % p
        f = F.mul(F.sqr(f), line(tx, ty, lam))
        nx = (lam * lam - 2 * tx) % p
        ty = (lam * (tx - nx) - ty) % p
        tx = nx
        if bit == "1":
            if tx == px:                      # T = -P: the chord is vertical (only at the very last step)
                tx = ty = None
                continue
            lam = (py - ty) * pow(px - tx, -1, p) % p
            f = F.mul(f, line(tx, ty, lam))
            nx = (lam * lam - tx - px) % p
            ty = (lam * (tx - nx) - ty) % p
            tx = nx
    return f


def final_exponentiation(P: PairingParams, F: Fq12, f):
    return F.pow(f, (F.p ** 12 - 1) // P.g1.fr.p)


def pairing(name: str, P1, Q2) -> List[int]:
    P = PAIRINGS[name]
    F = Fq12(P)
    return final_exponentiation(P, F, miller_loop(P, F, P1, Q2))


def product_of_pairings_is_one(name: str, pairs) -> bool:
    """prod e(P_i, Q_i) == 1 with one shared final exponentiation (``product_of_pairings``)."""
    P = PAIRINGS[name]
    F = Fq12(P)
    f = F.one
    for P1, Q2 in pairs:
        f = F.mul(f, miller_loop(P, F, P1, Q2))
    return final_exponentiation(P, F, f) == F.one


def groth16_verify(name: str, alpha_g1, beta_g2, gamma_g2, delta_g2, gamma_abc_g1, public_inputs: Sequence[int], proof) -> bool:
    """``verify_with_processed_vk``: e(A, B) = e(alpha, beta) e(IC, gamma) e(C, delta) with
    IC = gamma_abc[0] + sum x_i gamma_abc[i + 1], checked as a product equal to one."""
    P = PAIRINGS[name]
    g1 = P.g1
    a, b, c = proof
    ic = g1.to_jac(gamma_abc_g1[0])
    for x, pt in zip(public_inputs, gamma_abc_g1[1:]):
        ic = g1.add_jac(ic, g1.mul_scalar(pt, x % g1.fr.p))
    ic = g1.to_affine(ic)
    return product_of_pairings_is_one(name, [(g1.neg(a), b), (alpha_g1, beta_g2), (ic, gamma_g2), (c, delta_g2)])
