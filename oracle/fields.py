"""Prime fields used on the hot path (oracle; test infrastructure only).

Restates the parameters of ``ark_ff::Fp256<FrParameters>`` / ``Fp384<FqParameters>`` from
ark-bls12-381 / ark-bn254 0.3.0 (reached through ``pub mod ff`` / ``pub use bls12_381`` /
``pub use bn254`` in ``/root/reference/plugins/arkworks/src/lib.rs:16-20,101-103``).
Memory convention of those types: little-endian u64 limbs holding the MONTGOMERY residue
``a*R mod p`` with ``R = 2^(64*limbs)``; ``into_repr()`` yields the canonical integer.
"""
from __future__ import annotations

from dataclasses import dataclass


@dataclass(frozen=True)
class Field:
    name: str
    p: int
    limbs64: int          # number of u64 limbs (ark BigInteger width)
    generator: int = 0    # multiplicative generator (scalar fields only)
    two_adicity: int = 0  # scalar fields only

    # -- derived Montgomery constants ---------------------------------------------------
    @property
    def bits(self) -> int:
        return self.p.bit_length()

    @property
    def R(self) -> int:
        return (1 << (64 * self.limbs64)) % self.p

    @property
    def R2(self) -> int:
        return (self.R * self.R) % self.p

    @property
    def inv64(self) -> int:
        """-p^{-1} mod 2^64 (ark's ``INV``)."""
        return (-pow(self.p, -1, 1 << 64)) % (1 << 64)

    @property
    def inv32(self) -> int:
        return self.inv64 & 0xFFFFFFFF

    # -- arithmetic ---------------------------------------------------------------------
    def add(self, a, b):
        return (a + b) % self.p

    def sub(self, a, b):
        return (a - b) % self.p

    def mul(self, a, b):
        return (a * b) % self.p

    def neg(self, a):
        return (-a) % self.p

    def inv(self, a):
        if a % self.p == 0:
            raise ZeroDivisionError("inverse of zero")
        return pow(a, -1, self.p)

    def pow(self, a, e):
        return pow(a, e, self.p)

    # -- Montgomery (the in-memory form) ------------------------------------------------
    def to_mont(self, a: int) -> int:
        return (a * self.R) % self.p

    def from_mont(self, a: int) -> int:
        return (a * pow(self.R, -1, self.p)) % self.p

    def mont_mul(self, a: int, b: int) -> int:
        """Montgomery product of two Montgomery residues: a*b*R^-1 mod p."""
        return (a * b * pow(self.R, -1, self.p)) % self.p

    # -- limb packing -------------------------------------------------------------------
    def to_limbs(self, a: int):
        return [(a >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(self.limbs64)]

    @staticmethod
    def from_limbs(limbs) -> int:
        v = 0
        for i, l in enumerate(limbs):
            v |= int(l) << (64 * i)
        return v

    # -- scalar-field helpers (ark FftParameters) ---------------------------------------
    def two_adic_root(self) -> int:
        """``TWO_ADIC_ROOT_OF_UNITY = generator^((p-1)/2^TWO_ADICITY)``."""
        return pow(self.generator, (self.p - 1) >> self.two_adicity, self.p)

    def root_of_unity(self, log_n: int) -> int:
        """omega for a size-2^log_n Radix2EvaluationDomain (ark ``get_root_of_unity``)."""
        if log_n > self.two_adicity:
            raise ValueError("domain too large for the field's two-adicity")
        w = self.two_adic_root()
        for _ in range(self.two_adicity - log_n):
            w = (w * w) % self.p
        return w


BLS12_381_FQ = Field(
    "bls12_381_fq",
    0x1A0111EA397FE69A4B1BA7B6434BACD764774B84F38512BF6730D2A0F6B0F6241EABFFFEB153FFFFB9FEFFFFFFFFAAAB,
    6,
)
BLS12_381_FR = Field(
    "bls12_381_fr",
    0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001,
    4,
    generator=7,
    two_adicity=32,
)
BN254_FQ = Field(
    "bn254_fq",
    21888242871839275222246405745257275088696311157297823662689037894645226208583,
    4,
)
BN254_FR = Field(
    "bn254_fr",
    21888242871839275222246405745257275088548364400416034343698204186575808495617,
    4,
    generator=5,
    two_adicity=28,
)

FIELDS = {f.name: f for f in (BLS12_381_FQ, BLS12_381_FR, BN254_FQ, BN254_FR)}


def self_check() -> None:
    """Numerical cross-checks of the recalled ark constants."""
    assert BLS12_381_FQ.bits == 381 and BLS12_381_FR.bits == 255
    assert BN254_FQ.bits == 254 and BN254_FR.bits == 254
    assert BLS12_381_FQ.inv64 == 0x89F3FFFCFFFCFFFD
    assert BLS12_381_FR.inv64 == 0xFFFFFFFEFFFFFFFF
    assert BN254_FQ.inv64 == 0x87D20782E4866389
    assert BN254_FR.inv64 == 0xC2E1F593EFFFFFFF
    for f in (BLS12_381_FR, BN254_FR):
        assert (f.p - 1) % (1 << f.two_adicity) == 0
        assert ((f.p - 1) >> f.two_adicity) % 2 == 1
        w = f.two_adic_root()
        assert pow(w, 1 << f.two_adicity, f.p) == 1
        assert pow(w, 1 << (f.two_adicity - 1), f.p) == f.p - 1
    # Montgomery limbs of the two-adic roots as recalled from the ark parameter files.
    assert BLS12_381_FR.to_limbs(BLS12_381_FR.to_mont(BLS12_381_FR.two_adic_root())) == [
        0xB9B58D8C5F0E466A, 0x5B1B4C801819D7EC, 0x0AF53AE352A31E64, 0x5BF3ADDA19E9B27B]
    assert BN254_FR.to_limbs(BN254_FR.to_mont(BN254_FR.two_adic_root())) == [
        7164790868263648668, 11685701338293206998, 6216421865291908056, 1756667274303109607]
