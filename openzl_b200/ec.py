"""Mirror of ``ark_ec::msm::VariableBaseMSM`` (ark-ec 0.3.0) as re-exported by
``plugins/arkworks`` (`pub use ec`, /root/reference/plugins/arkworks/src/lib.rs:28-29).

ark signature::

    VariableBaseMSM::multi_scalar_mul(bases: &[G], scalars: &[<G::ScalarField as PrimeField>::BigInt]) -> G::Projective

Here ``bases`` is an ``(n, 2*L)`` uint64 array of packed affine points (Montgomery limbs, the
in-memory form of ``GroupAffine``; optional ``infinity`` bitset) or a device-resident
:class:`openzl_b200.Bases`; ``scalars`` an ``(n, 4)`` uint64 array of canonical ``BigInteger256``;
the result a :class:`GroupProjective` holding Jacobian ``X||Y||Z`` Montgomery limbs.
As in ark, ``size = min(len(bases), len(scalars))``.
"""
from __future__ import annotations

from typing import Optional, Union

import numpy as np

from . import _lib
from .context import Bases, Context


class GroupProjective:
    """Jacobian point as returned by ``multi_scalar_mul`` (x = X/Z^2, y = Y/Z^3)."""

    def __init__(self, curve: int, limbs: np.ndarray, ctx: Context):
        self.curve, self.limbs, self._ctx = curve, limbs, ctx

    def is_zero(self) -> bool:
        n = len(self.limbs) // 3
        return not self.limbs[2 * n:].any()

    def into_affine(self):
        """``(x||y limbs, infinity)`` -- ``GroupProjective::into_affine``; runs on the device."""
        return self._ctx.jacobian_to_affine(self.curve, self.limbs)


class VariableBaseMSM:
    @staticmethod
    def multi_scalar_mul(bases: Union[np.ndarray, Bases], scalars: np.ndarray, *, curve: Optional[int] = None,
                         infinity: Optional[np.ndarray] = None, ctx: Optional[Context] = None) -> GroupProjective:
        from . import default_context
        scalars = np.ascontiguousarray(scalars, dtype=np.uint64)
        if isinstance(bases, Bases):
            ctx = bases.ctx
            size = min(bases.n, scalars.shape[0])
            return GroupProjective(bases.curve, bases.msm(scalars[:size]), ctx)
        if curve is None:
            raise _lib.OzlError(1, "multi_scalar_mul", "curve= is required when bases is a host array")
        ctx = ctx or default_context()
        size = min(bases.shape[0], scalars.shape[0])
        mask = None
        if infinity is not None:
            mask = np.ascontiguousarray(infinity, dtype=np.uint8)
        handle = ctx.upload_bases(curve, bases[:size], mask)
        try:
            return GroupProjective(curve, handle.msm(scalars[:size]), ctx)
        finally:
            handle.free()
