"""openzl_b200 -- Blackwell-native backend for the two compute kernels under OpenZL's
``Groth16::prove`` (/root/reference/plugins/arkworks/src/groth16.rs:445-457): Pippenger MSM and
radix-2 NTT, as hand-written sm_100a CUDA behind the C ABI in ``include/ozl.h``.

Host-side mirror of the reference's surface (same names and argument meaning):

* ``openzl_b200.ec.VariableBaseMSM.multi_scalar_mul(bases, scalars)``  -- ark_ec::msm, via `pub use ec`
* ``openzl_b200.poly.Radix2EvaluationDomain`` / ``GeneralEvaluationDomain`` -- ark_poly, via `pub use poly`
* ``openzl_b200.groth16.Groth16.{compile, prove}`` -- the plugin's ``ProofSystem`` impl (groth16.rs:399-467)
* ``openzl_b200.serialize`` -- ark ``CanonicalSerialize`` bytes of ``Proof`` / ``ProvingKey`` (groth16.rs:98-179)

Importing this package never touches ``oracle/`` and there is no CPU fallback: constructing a
``Context`` without ``libozl_b200.so`` or without a CUDA device raises.
"""
from ._lib import (BLS12_381_FR, BLS12_381_G1, BLS12_381_G2, BN254_FR, BN254_G1, BN254_G2, CURVE_IDS, FIELD_IDS,
                   OzlError, OzlLibraryError)
from .context import Bases, Context
from . import ec, poly, serialize

__all__ = ["Context", "Bases", "OzlError", "OzlLibraryError", "ec", "poly", "serialize", "BLS12_381_G1", "BLS12_381_G2", "BN254_G1",
           "BN254_G2", "BN254_FR", "BLS12_381_FR", "CURVE_IDS", "FIELD_IDS", "default_context"]

_default_ctx = None


def default_context(device: int = 0) -> Context:
    """Process-wide context used by the static-method style API (ark's associated functions
    carry no state: ``Groth16<E>(PhantomData)``, groth16.rs:399-403)."""
    global _default_ctx
    if _default_ctx is None or _default_ctx._h is None or _default_ctx.device != device:
        _default_ctx = Context(device)
    return _default_ctx
