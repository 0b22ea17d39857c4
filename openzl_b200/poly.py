"""Mirror of ``ark_poly::{EvaluationDomain, Radix2EvaluationDomain, GeneralEvaluationDomain}``
(ark-poly 0.3.0) as re-exported by ``plugins/arkworks`` (`pub use poly`,
/root/reference/plugins/arkworks/src/lib.rs:70-71) -- the methods
``ark_groth16::R1CStoQAP::witness_map`` uses.  Vectors are ``(len, 4)`` uint64 arrays of
Montgomery limbs (the in-memory form of ``Fp256``); like ark, inputs shorter than the domain are
zero-extended and the ``*_in_place`` forms resize to ``size()``.
"""
from __future__ import annotations

from typing import Optional

import numpy as np

from . import _lib
from .context import Context

_TWO_ADICITY = {_lib.BN254_FR: 28, _lib.BLS12_381_FR: 32}
# ark `FftParameters` / `FpParameters` of the two scalar fields: modulus and multiplicative generator
# (TWO_ADIC_ROOT_OF_UNITY = GENERATOR^((r - 1) / 2^TWO_ADICITY)); the device NTT uses the same constants.
_MODULUS = {_lib.BN254_FR: 21888242871839275222246405745257275088548364400416034343698204186575808495617,
            _lib.BLS12_381_FR: 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001}
_GENERATOR = {_lib.BN254_FR: 5, _lib.BLS12_381_FR: 7}


class Radix2EvaluationDomain:
    def __init__(self, field: int, size: int, log_size_of_group: int, ctx: Context):
        self.field, self._size, self.log_size_of_group, self._ctx = field, size, log_size_of_group, ctx

    @classmethod
    def new(cls, field: int, num_coeffs: int, ctx: Optional[Context] = None) -> Optional["Radix2EvaluationDomain"]:
        """Smallest power-of-two domain holding ``num_coeffs``; ``None`` when it exceeds the
        field's two-adicity (ark: ``Radix2EvaluationDomain::new`` returns ``None``)."""
        from . import default_context
        if field not in _TWO_ADICITY:
            raise _lib.OzlError(1, "Radix2EvaluationDomain.new", "unknown field")
        size, log = 1, 0
        while size < num_coeffs:
            size <<= 1
            log += 1
        if log > _TWO_ADICITY[field]:
            return None
        return cls(field, size, log, ctx or default_context())

    def size(self) -> int:
        return self._size

    # ---- domain constants and the O(1) / O(n) host-side helpers of the trait ----------------
    # Returned as canonical Python integers (``Fr::into_repr()``); they parameterise setup-time
    # formulas, not the per-proof hot path.
    @property
    def modulus(self) -> int:
        return _MODULUS[self.field]

    @property
    def group_gen(self) -> int:
        """omega: ``TWO_ADIC_ROOT_OF_UNITY ^ (2^(TWO_ADICITY - log_size_of_group))``."""
        p = self.modulus
        w = pow(_GENERATOR[self.field], (p - 1) >> _TWO_ADICITY[self.field], p)
        for _ in range(_TWO_ADICITY[self.field] - self.log_size_of_group):
            w = (w * w) % p
        return w

    @property
    def group_gen_inv(self) -> int:
        return pow(self.group_gen, -1, self.modulus)

    @property
    def size_inv(self) -> int:
        return pow(self._size, -1, self.modulus)

    @property
    def generator_inv(self) -> int:
        """Inverse of ``F::multiplicative_generator()``, the coset shift."""
        return pow(_GENERATOR[self.field], -1, self.modulus)

    def element(self, i: int) -> int:
        """``domain.element(i)`` = omega^i."""
        return pow(self.group_gen, i, self.modulus)

    def elements(self):
        p, w, cur = self.modulus, self.group_gen, 1
        for _ in range(self._size):
            yield cur
            cur = (cur * w) % p

    def evaluate_vanishing_polynomial(self, tau: int) -> int:
        """Z(tau) = tau^size - 1."""
        return (pow(tau, self._size, self.modulus) - 1) % self.modulus

    def evaluate_all_lagrange_coefficients(self, tau: int):
        """L_i(tau) for all i: Z(tau) * omega^i / (size * (tau - omega^i)); when tau is in the
        domain the indicator vector, as ark returns."""
        p, n = self.modulus, self._size
        z = self.evaluate_vanishing_polynomial(tau)
        if z == 0:
            return [1 if e == tau % p else 0 for e in self.elements()]
        # one batched inversion (Montgomery's trick) for the n denominators
        dens = [(n * (tau - e)) % p for e in self.elements()]
        pref, acc = [], 1
        for d in dens:
            pref.append(acc)
            acc = (acc * d) % p
        inv = pow(acc, -1, p)
        out = [0] * n
        for i in range(n - 1, -1, -1):
            out[i] = (inv * pref[i]) % p
            inv = (inv * dens[i]) % p
        cur, w = 1, self.group_gen
        for i in range(n):
            out[i] = (out[i] * z % p) * cur % p
            cur = (cur * w) % p
        return out

    def _prep(self, a: np.ndarray) -> np.ndarray:
        a = np.asarray(a, dtype=np.uint64)
        if a.ndim != 2 or a.shape[1] != 4:
            raise _lib.OzlError(1, "EvaluationDomain", "expected (len, 4) uint64 Montgomery limbs")
        if a.shape[0] > self._size:
            raise _lib.OzlError(1, "EvaluationDomain", "input longer than the domain")
        out = np.zeros((self._size, 4), dtype=np.uint64)
        out[: a.shape[0]] = a
        return out

    def _run(self, a, inverse, coset):
        buf = self._prep(a)
        self._ctx.ntt(self.field, buf, inverse=inverse, coset=coset)
        return buf

    def fft(self, coeffs: np.ndarray) -> np.ndarray:
        return self._run(coeffs, False, False)

    def ifft(self, evals: np.ndarray) -> np.ndarray:
        return self._run(evals, True, False)

    def coset_fft(self, coeffs: np.ndarray) -> np.ndarray:
        return self._run(coeffs, False, True)

    def coset_ifft(self, evals: np.ndarray) -> np.ndarray:
        return self._run(evals, True, True)

    # ``*_in_place`` take a list-like holder ``[array]`` semantics in Rust (``&mut Vec<F>``);
    # numpy arrays cannot be resized in place, so these return the (possibly longer) array.
    fft_in_place = fft
    ifft_in_place = ifft
    coset_fft_in_place = coset_fft
    coset_ifft_in_place = coset_ifft


class GeneralEvaluationDomain(Radix2EvaluationDomain):
    """ark's enum picks Radix2 whenever a power-of-two domain fits (always for BN254/BLS12-381 Fr
    up to 2^28 / 2^32; their mixed-radix branch is not configured for these fields)."""
