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
