"""Short-Weierstrass groups on the hot path (oracle; test infrastructure only).

Restates ``ark_ec::models::short_weierstrass_jacobian::{GroupAffine, GroupProjective}``
(ark-ec 0.3.0) for the four groups behind ``Pairing::{G1, G2}``
(``/root/reference/plugins/arkworks/src/pairing.rs:14-23``): BLS12-381 G1/G2, BN254 G1/G2,
all with a = 0.  Points are Python tuples of canonical integers:

* affine   ``(x, y)`` or ``None`` for the point at infinity (ark: ``infinity = true``)
* Jacobian ``(X, Y, Z)`` with x = X/Z^2, y = Y/Z^3, identity ``(0, 1, 0)`` (any Z == 0)

For G2 every coordinate is an Fq2 element ``(c0, c1)`` = c0 + c1*u with u^2 = -1.
Formulas: madd-2007-bl, add-2007-bl, dbl-2009-l -- the ones ark's
``add_assign_mixed`` / ``add_assign`` / ``double_in_place`` use.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any

from .fields import BLS12_381_FQ, BLS12_381_FR, BN254_FQ, BN254_FR, Field


class FpOps:
    """Field-operation adaptor over a prime field (elements are ints)."""

    def __init__(self, f: Field):
        self.f = f
        self.p = f.p
        self.zero = 0
        self.one = 1
        self.degree = 1

    def add(self, a, b): return (a + b) % self.p
    def sub(self, a, b): return (a - b) % self.p
    def mul(self, a, b): return (a * b) % self.p
    def sqr(self, a): return (a * a) % self.p
    def neg(self, a): return (-a) % self.p
    def inv(self, a): return pow(a, -1, self.p)
    def is_zero(self, a): return a % self.p == 0
    def from_int(self, v): return v % self.p
    def dbl(self, a): return (2 * a) % self.p
    def coords(self, a): return [a]
    def from_coords(self, c): return c[0]


class Fp2Ops:
    """Fq2 = Fq[u]/(u^2+1); elements are (c0, c1) tuples (ark ``QuadExtField``)."""

    def __init__(self, f: Field):
        self.f = f
        self.p = f.p
        self.zero = (0, 0)
        self.one = (1, 0)
        self.degree = 2

    def add(self, a, b): return ((a[0] + b[0]) % self.p, (a[1] + b[1]) % self.p)
    def sub(self, a, b): return ((a[0] - b[0]) % self.p, (a[1] - b[1]) % self.p)
    def mul(self, a, b):
        p = self.p
        return ((a[0] * b[0] - a[1] * b[1]) % p, (a[0] * b[1] + a[1] * b[0]) % p)
    def sqr(self, a): return self.mul(a, a)
    def neg(self, a): return ((-a[0]) % self.p, (-a[1]) % self.p)
    def inv(self, a):
        p = self.p
        n = pow(a[0] * a[0] + a[1] * a[1], -1, p)
        return ((a[0] * n) % p, (-a[1] * n) % p)
    def is_zero(self, a): return a[0] % self.p == 0 and a[1] % self.p == 0
    def from_int(self, v): return (v % self.p, 0)
    def dbl(self, a): return ((2 * a[0]) % self.p, (2 * a[1]) % self.p)
    def coords(self, a): return [a[0], a[1]]
    def from_coords(self, c): return (c[0], c[1])


@dataclass
class Curve:
    name: str
    F: Any               # FpOps | Fp2Ops
    b: Any               # curve coefficient (a = 0)
    gen: tuple           # affine generator
    fr: Field            # scalar field
    base: Field          # base prime field

    # ---- predicates / conversions ------------------------------------------------------
    def is_on_curve(self, P) -> bool:
        if P is None:
            return True
        F = self.F
        x, y = P
        return F.sub(F.sqr(y), F.add(F.mul(F.sqr(x), x), self.b)) == F.zero

    def neg(self, P):
        return None if P is None else (P[0], self.F.neg(P[1]))

    def identity_jac(self):
        return (self.F.zero, self.F.one, self.F.zero)  # ark GroupProjective::zero()

    def to_jac(self, P):
        return self.identity_jac() if P is None else (P[0], P[1], self.F.one)

    def to_affine(self, J):
        F = self.F
        X, Y, Z = J
        if F.is_zero(Z):
            return None
        zi = F.inv(Z)
        zi2 = F.sqr(zi)
        return (F.mul(X, zi2), F.mul(Y, F.mul(zi2, zi)))

    # ---- affine group law (independent path used to validate the Jacobian formulas) ----
    def add_affine(self, P, Q):
        F = self.F
        if P is None:
            return Q
        if Q is None:
            return P
        if P[0] == Q[0]:
            if P[1] == Q[1] and not F.is_zero(P[1]):
                lam = F.mul(F.mul(F.from_int(3), F.sqr(P[0])), F.inv(F.dbl(P[1])))
            else:
                return None
        else:
            lam = F.mul(F.sub(Q[1], P[1]), F.inv(F.sub(Q[0], P[0])))
        x3 = F.sub(F.sub(F.sqr(lam), P[0]), Q[0])
        y3 = F.sub(F.mul(lam, F.sub(P[0], x3)), P[1])
        return (x3, y3)

    # ---- Jacobian formulas, a = 0 ------------------------------------------------------
    def dbl_jac(self, J):
        """dbl-2009-l (ark ``double_in_place`` for a = 0)."""
        F = self.F
        X1, Y1, Z1 = J
        if F.is_zero(Z1):
            return J
        A = F.sqr(X1)
        B = F.sqr(Y1)
        C = F.sqr(B)
        D = F.dbl(F.sub(F.sub(F.sqr(F.add(X1, B)), A), C))
        E = F.add(F.dbl(A), A)
        Fq = F.sqr(E)
        X3 = F.sub(Fq, F.dbl(D))
        Y3 = F.sub(F.mul(E, F.sub(D, X3)), F.dbl(F.dbl(F.dbl(C))))
        Z3 = F.dbl(F.mul(Y1, Z1))
        return (X3, Y3, Z3)

    def add_mixed(self, J, P):
        """madd-2007-bl with ark's guards (``add_assign_mixed``)."""
        F = self.F
        if P is None:
            return J
        X1, Y1, Z1 = J
        if F.is_zero(Z1):
            return (P[0], P[1], F.one)
        Z1Z1 = F.sqr(Z1)
        U2 = F.mul(P[0], Z1Z1)
        S2 = F.mul(F.mul(P[1], Z1), Z1Z1)
        if X1 == U2 and Y1 == S2:
            return self.dbl_jac(J)
        H = F.sub(U2, X1)
        HH = F.sqr(H)
        I = F.dbl(F.dbl(HH))
        Jv = F.mul(H, I)
        r = F.dbl(F.sub(S2, Y1))
        V = F.mul(X1, I)
        X3 = F.sub(F.sub(F.sqr(r), Jv), F.dbl(V))
        Y3 = F.sub(F.mul(r, F.sub(V, X3)), F.dbl(F.mul(Y1, Jv)))
        Z3 = F.sub(F.sub(F.sqr(F.add(Z1, H)), Z1Z1), HH)
        return (X3, Y3, Z3)

    def add_jac(self, J1, J2):
        """add-2007-bl with ark's guards (``AddAssign``)."""
        F = self.F
        X1, Y1, Z1 = J1
        X2, Y2, Z2 = J2
        if F.is_zero(Z1):
            return J2
        if F.is_zero(Z2):
            return J1
        Z1Z1 = F.sqr(Z1)
        Z2Z2 = F.sqr(Z2)
        U1 = F.mul(X1, Z2Z2)
        U2 = F.mul(X2, Z1Z1)
        S1 = F.mul(F.mul(Y1, Z2), Z2Z2)
        S2 = F.mul(F.mul(Y2, Z1), Z1Z1)
        if U1 == U2 and S1 == S2:
            return self.dbl_jac(J1)
        H = F.sub(U2, U1)
        I = F.sqr(F.dbl(H))
        Jv = F.mul(H, I)
        r = F.dbl(F.sub(S2, S1))
        V = F.mul(U1, I)
        X3 = F.sub(F.sub(F.sqr(r), Jv), F.dbl(V))
        Y3 = F.sub(F.mul(r, F.sub(V, X3)), F.dbl(F.mul(S1, Jv)))
        Z3 = F.mul(F.sub(F.sub(F.sqr(F.add(Z1, Z2)), Z1Z1), Z2Z2), H)
        return (X3, Y3, Z3)

    def mul_scalar(self, P, k: int):
        """[k]P for an affine P by left-to-right double-and-add (ark ``AffineCurve::mul``)."""
        acc = self.identity_jac()
        if P is None or k == 0:
            return acc
        if k < 0:
            return self.mul_scalar(self.neg(P), -k)
        for bit in bin(k)[2:]:
            acc = self.dbl_jac(acc)
            if bit == "1":
                acc = self.add_mixed(acc, P)
        return acc

    def mul_affine(self, P, k: int):
        return self.to_affine(self.mul_scalar(P, k))

    # ---- in-memory layout (what the C ABI takes) ---------------------------------------
    def coord_limbs(self) -> int:
        return self.base.limbs64 * self.F.degree

    def affine_to_mont_limbs(self, P):
        """Packed affine point: x||y (G2: x.c0||x.c1||y.c0||y.c1), u64 LE limbs, Montgomery.
        The point at infinity is encoded as all-zero limbs (flagged separately)."""
        n = self.base.limbs64
        if P is None:
            return [0] * (2 * n * self.F.degree)
        out = []
        for c in self.F.coords(P[0]) + self.F.coords(P[1]):
            out += self.base.to_limbs(self.base.to_mont(c))
        return out

    def jac_from_mont_limbs(self, limbs):
        """Inverse of the C ABI's output layout X||Y||Z (Montgomery u64 limbs)."""
        n = self.base.limbs64
        d = self.F.degree
        vals = [self.base.from_mont(Field.from_limbs(limbs[i * n:(i + 1) * n])) for i in range(3 * d)]
        cs = [self.F.from_coords(vals[i * d:(i + 1) * d]) for i in range(3)]
        return tuple(cs)

    def affine_from_mont_limbs(self, limbs):
        n = self.base.limbs64
        d = self.F.degree
        vals = [self.base.from_mont(Field.from_limbs(limbs[i * n:(i + 1) * n])) for i in range(2 * d)]
        return (self.F.from_coords(vals[0:d]), self.F.from_coords(vals[d:2 * d]))


_fq381 = FpOps(BLS12_381_FQ)
_fq381_2 = Fp2Ops(BLS12_381_FQ)
_fq254 = FpOps(BN254_FQ)
_fq254_2 = Fp2Ops(BN254_FQ)

BLS12_381_G1 = Curve(
    "bls12_381_g1", _fq381, 4,
    (3685416753713387016781088315183077757961620795782546409894578378688607592378376318836054947676345821548104185464507,
     1339506544944476473020471379941921221584933875938349620426543736416511423956333506472724655353366534992391756441569),
    BLS12_381_FR, BLS12_381_FQ)

BLS12_381_G2 = Curve(
    "bls12_381_g2", _fq381_2, (4, 4),
    ((352701069587466618187139116011060144890029952792775240219908644239793785735715026873347600343865175952761926303160,
      3059144344244213709971259814753781636986470325476647558659373206291635324768958432433509563104347017837885763365758),
     (1985150602287291935568054521177171638300868978215655730859378665066344726373823718423869104263333984641494340347905,
      927553665492332455747201965776037880757740193453592970025027978793976877002675564980949289727957565575433344219582)),
    BLS12_381_FR, BLS12_381_FQ)

BN254_G1 = Curve("bn254_g1", _fq254, 3, (1, 2), BN254_FR, BN254_FQ)

BN254_G2 = Curve(
    "bn254_g2", _fq254_2,
    (19485874751759354771024239261021720505790618469301721065564631296452457478373,
     266929791119991161246907387137283842545076965332900288569378510910307636690),
    ((10857046999023057135944570762232829481370756359578518086990519993285655852781,
      11559732032986387107991004021392285783925812861821192530917403151452391805634),
     (8495653923123431417604973247489272438418190587263600148770280649306958101930,
      4082367875863433681332203403145435568316851327593401208105741076214120093531)),
    BN254_FR, BN254_FQ)

CURVES = {c.name: c for c in (BLS12_381_G1, BLS12_381_G2, BN254_G1, BN254_G2)}


def self_check() -> None:
    for c in CURVES.values():
        assert c.is_on_curve(c.gen), c.name
        # the generator has prime order r (so [r]G = identity)
        assert c.to_affine(c.mul_scalar(c.gen, c.fr.p)) is None, c.name
        two = c.mul_affine(c.gen, 2)
        assert two == c.add_affine(c.gen, c.gen)
        assert c.mul_affine(c.gen, 5) == c.add_affine(c.add_affine(two, two), c.gen)
