"""Multi-scalar multiplication (oracle; test infrastructure only).

``msm_ark`` restates ``ark_ec::msm::VariableBaseMSM::multi_scalar_mul`` (ark-ec 0.3.0,
``src/msm/variable_base.rs``; reached via ``pub use ec`` in
``/root/reference/plugins/arkworks/src/lib.rs:28-29``; called by ``ark_groth16::create_proof``
behind ``/root/reference/plugins/arkworks/src/groth16.rs:454``) step for step: window rule,
unsigned c-bit digits, zero-scalar filter, ``scalar == 1`` shortcut in window 0,
``2^c - 1`` Jacobian buckets, top-down running sum, high-to-low fold with c doublings.
PARITY UNPINNED by the reference's tests; validated against ``msm_naive`` and the
known-discrete-log identity.
"""
from __future__ import annotations

import math

from .curves import Curve


def ark_window_bits(size: int) -> int:
    """``c`` as ark-ec 0.3.0 picks it: 3 if size < 32 else ln_without_floats(size) + 2,
    with ``ln_without_floats(a) = ceil(log2(a)) * 69 / 100`` (integer division)."""
    if size < 32:
        return 3
    log2 = (size - 1).bit_length() if size > 1 else 0  # ceil(log2(size))
    return (log2 * 69) // 100 + 2


def msm_naive(curve: Curve, bases, scalars):
    """sum_i [s_i]P_i by independent double-and-add (Jacobian result)."""
    acc = curve.identity_jac()
    for P, s in zip(bases, scalars):
        acc = curve.add_jac(acc, curve.mul_scalar(P, s))
    return acc


def msm_ark(curve: Curve, bases, scalars):
    """ark-ec 0.3.0 ``VariableBaseMSM::multi_scalar_mul`` (serial path). Returns Jacobian."""
    size = min(len(bases), len(scalars))
    pairs = [(scalars[i], bases[i]) for i in range(size) if scalars[i] != 0]
    c = ark_window_bits(size)
    num_bits = curve.fr.bits
    zero = curve.identity_jac()
    window_sums = []
    for w_start in range(0, num_bits, c):
        res = zero
        buckets = [zero] * ((1 << c) - 1)
        for s, P in pairs:
            if s == 1:
                if w_start == 0:
                    res = curve.add_mixed(res, P)
            else:
                d = (s >> w_start) % (1 << c)
                if d != 0:
                    buckets[d - 1] = curve.add_mixed(buckets[d - 1], P)
        running = zero
        for b in reversed(buckets):
            running = curve.add_jac(running, b)
            res = curve.add_jac(res, running)
        window_sums.append(res)
    lowest = window_sums[0]
    total = zero
    for ws in reversed(window_sums[1:]):
        total = curve.add_jac(total, ws)
        for _ in range(c):
            total = curve.dbl_jac(total)
    return curve.add_jac(lowest, total)


def msm_known_dlog(curve: Curve, dlogs, scalars):
    """For bases P_i = [d_i]G: sum s_i P_i = [sum s_i d_i mod r]G.  Works at any n."""
    r = curve.fr.p
    k = 0
    for d, s in zip(dlogs, scalars):
        k = (k + d * s) % r
    return curve.mul_scalar(curve.gen, k)
