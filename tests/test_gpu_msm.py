"""-m gpu parity tests of the MSM path, through the C ABI, against the CPU oracle.

Contract under test: ark_ec::msm::VariableBaseMSM::multi_scalar_mul as called behind
/root/reference/plugins/arkworks/src/groth16.rs:454.  Bit-exact: affine (x, y) Montgomery limbs
of the GPU result equal the oracle's (Jacobian representatives are not unique, so comparison is
after into_affine, as SURVEY.md section 7 prescribes).
"""
import numpy as np
import pytest

import openzl_b200 as ozl
from oracle import cbind, curves
from tests.util import ints_to_array, limbs_to_int, random_scalars

pytestmark = pytest.mark.gpu

CURVES = ["bls12_381_g1", "bn254_g1", "bls12_381_g2", "bn254_g2"]


def oracle_affine(name, bases, scalars, inf=None, threads=8):
    return cbind.to_affine(name, cbind.msm(name, bases, scalars, inf=inf, threads=threads))


def gpu_affine(ctx, name, jac):
    aff, is_inf = ctx.jacobian_to_affine(ozl.CURVE_IDS[name], jac)
    # cross-check the device normalisation with the oracle's
    o_aff, o_inf = cbind.to_affine(name, jac)
    assert is_inf == o_inf
    assert (aff == o_aff).all()
    return aff, is_inf


def directed_scalars(name, n, seed):
    c = curves.CURVES[name]
    r = c.fr.p
    s = random_scalars(n, r, seed)
    special = [0, 1, 2, r - 1, r - 2, (1 << 16) - 1, 1 << 16, (1 << 15), (1 << 15) + 1, (1 << 254) - 1 if r > (1 << 254) else (1 << 253) - 1,
               (r - 1) // 2, (r + 1) // 2, 0xFFFF_FFFF, 1 << 32, (1 << 64) - 1, 1 << 64]
    for i, v in enumerate(special):
        s[i] = ints_to_array([v % r])[0]
    return s


@pytest.mark.parametrize("name", CURVES)
def test_msm_2_12_parity(ctx, name):
    """BASELINE config 1: 2^12 points, seeded random + directed scalars, directed points."""
    n = 1 << 12 if name.endswith("g1") else 1 << 10
    bases = cbind.bases_seq(name, 1, n)
    # directed points: a duplicate, P and -P, and (below) an infinity entry
    bases[5] = bases[4]
    c = curves.CURVES[name]
    L = cbind.COORD_LIMBS[name]
    P7 = c.affine_from_mont_limbs(list(bases[7]))
    bases[8] = np.array(c.affine_to_mont_limbs(c.neg(P7)), dtype=np.uint64)
    scalars = directed_scalars(name, n, seed=0x4F5A4C5F)
    scalars[5] = scalars[4]          # P + P inside one bucket -> doubling branch
    scalars[8] = scalars[7]          # P + (-P) inside one bucket -> identity branch
    inf = np.zeros((n + 7) // 8, dtype=np.uint8)
    inf[20 >> 3] |= 1 << (20 & 7)
    exp, exp_inf = oracle_affine(name, bases, scalars, inf=inf)
    h = ctx.upload_bases(ozl.CURVE_IDS[name], bases, inf)
    try:
        got, got_inf = gpu_affine(ctx, name, h.msm(scalars))
    finally:
        h.free()
    assert got_inf == exp_inf
    assert (got == exp).all()


@pytest.mark.parametrize("c_bits", [4, 7, 10, 13, 16])
def test_msm_window_sweep(ctx, c_bits):
    name = "bls12_381_g1"
    n = 3000
    bases = cbind.bases_seq(name, 3, n)
    scalars = directed_scalars(name, n, seed=c_bits)
    exp, _ = oracle_affine(name, bases, scalars)
    h = ctx.upload_bases(ozl.BLS12_381_G1, bases)
    try:
        ctx.set_window_bits(c_bits)
        got, _ = gpu_affine(ctx, name, h.msm(scalars))
    finally:
        ctx.set_window_bits(0)
        h.free()
    assert (got == exp).all()


@pytest.mark.parametrize("n", [0, 1, 2, 31, 32, 33, 257])
def test_msm_ragged_sizes(ctx, n):
    name = "bn254_g1"
    bases = cbind.bases_seq(name, 1, max(n, 1))
    scalars = random_scalars(max(n, 1), curves.CURVES[name].fr.p, seed=n + 1)
    h = ctx.upload_bases(ozl.BN254_G1, bases)
    try:
        jac = h.msm(scalars[:n])
    finally:
        h.free()
    got, got_inf = gpu_affine(ctx, name, jac)
    if n == 0:
        assert got_inf
        # ark GroupProjective::zero() = (0, 1, 0)
        assert limbs_to_int(jac[8:12]) == 0 and limbs_to_int(jac[0:4]) == 0
        return
    exp, exp_inf = oracle_affine(name, bases[:n], scalars[:n])
    assert got_inf == exp_inf and (got == exp).all()


def test_msm_all_zero_and_all_one_scalars(ctx):
    name = "bls12_381_g1"
    n = 5000
    bases = cbind.bases_seq(name, 1, n)
    h = ctx.upload_bases(ozl.BLS12_381_G1, bases)
    try:
        zero = np.zeros((n, 4), dtype=np.uint64)
        _, is_inf = gpu_affine(ctx, name, h.msm(zero))
        assert is_inf
        one = zero.copy()
        one[:, 0] = 1          # every point lands in one bucket: exercises the task splitter
        got, _ = gpu_affine(ctx, name, h.msm(one))
        exp, _ = oracle_affine(name, bases, one)
        assert (got == exp).all()
        # sum_{i<n} (i+1) G = [n(n+1)/2] G
        k = n * (n + 1) // 2
        assert (got == cbind.to_affine(name, cbind.gen_mul(name, k))[0]).all()
    finally:
        h.free()


@pytest.mark.parametrize("c_bits", [12, 14])
def test_msm_heavy_buckets(ctx, c_bits):
    """Window widths whose top window has only 2-3 live bits put ~n/4 points in a few buckets;
    with all-equal scalars every point of a window shares ONE bucket.  Both go through the
    heavy-bucket collapse passes (more than 8 partials per bucket)."""
    name = "bn254_g1"
    n = 40000
    r = curves.CURVES[name].fr.p
    bases = cbind.bases_seq(name, 1, n)
    h = ctx.upload_bases(ozl.BN254_G1, bases)
    try:
        ctx.set_window_bits(c_bits)
        s1 = random_scalars(n, r, seed=c_bits)
        got, _ = gpu_affine(ctx, name, h.msm(s1))
        assert (got == oracle_affine(name, bases, s1)[0]).all()
        s2 = np.tile(s1[:1], (n, 1))            # every scalar identical: one bucket per window
        got, _ = gpu_affine(ctx, name, h.msm(s2))
        assert (got == oracle_affine(name, bases, s2)[0]).all()
    finally:
        ctx.set_window_bits(0)
        h.free()


def test_msm_skewed_small_scalars(ctx):
    """Witness-like distribution: mostly 0/1/small values plus a few full-width ones."""
    name = "bn254_g1"
    n = 1 << 14
    r = curves.CURVES[name].fr.p
    rng = np.random.default_rng(7)
    scalars = np.zeros((n, 4), dtype=np.uint64)
    scalars[:, 0] = rng.integers(0, 3, size=n, dtype=np.uint64)
    full = random_scalars(n // 16, r, seed=8)
    scalars[:: 16] = full
    bases = cbind.bases_seq(name, 11, n)
    exp, _ = oracle_affine(name, bases, scalars)
    h = ctx.upload_bases(ozl.BN254_G1, bases)
    try:
        got, _ = gpu_affine(ctx, name, h.msm(scalars))
    finally:
        h.free()
    assert (got == exp).all()


@pytest.mark.parametrize("name", CURVES)
def test_generated_bases_match_oracle(ctx, name):
    n = 1000
    h = ctx.generate_bases(ozl.CURVE_IDS[name], 5, n)
    try:
        got = h.download()
    finally:
        h.free()
    exp = cbind.bases_seq(name, 5, n)
    assert (got == exp).all()


@pytest.mark.parametrize("log_n", [16, 20, 22, 24])
def test_msm_known_dlog_large(ctx, log_n):
    """Size-independent property: for P_i = [i+1]G, sum s_i P_i = [sum s_i (i+1) mod r] G.
    2^22 and 2^24 are BASELINE config 2's sweep sizes; from 2^22 up the host scalars (pageable numpy
    memory here) reach the device in point-range batches that overlap the accumulation."""
    name = "bls12_381_g1"
    n = 1 << log_n
    r = curves.CURVES[name].fr.p
    scalars = random_scalars(n, r, seed=log_n)
    h = ctx.generate_bases(ozl.BLS12_381_G1, 1, n)
    try:
        got, _ = gpu_affine(ctx, name, h.msm(scalars))
    finally:
        h.free()
    k = cbind.dot_mod_r("bls12_381_fr", scalars, np.arange(1, n + 1, dtype=np.uint64))
    exp, _ = cbind.to_affine(name, cbind.gen_mul(name, k))
    assert (got == exp).all()


@pytest.mark.parametrize("name,factor", [("bls12_381_g1", 2), ("bls12_381_g1", 4), ("bn254_g1", 3), ("bn254_g1", 8),
                                         ("bn254_g2", 4), ("bls12_381_g2", 2)])
def test_msm_precomputed_bases(ctx, name, factor):
    """ozl_msm_bases_precompute keeps the result bit-identical (shifted copies, fewer bucket sets)."""
    n = 3000 if name.endswith("g1") else 700
    bases = cbind.bases_seq(name, 2, n)
    scalars = directed_scalars(name, n, seed=factor)
    inf = np.zeros((n + 7) // 8, dtype=np.uint8)
    inf[3] = 0x10
    exp, _ = oracle_affine(name, bases, scalars, inf=inf)
    h = ctx.upload_bases(ozl.CURVE_IDS[name], bases, inf).precompute(factor)
    try:
        got, _ = gpu_affine(ctx, name, h.msm(scalars))
        assert (got == exp).all()
        # fewer scalars than bases on a precomputed handle (copies are strided by the handle's size)
        m = n // 3
        got2, _ = gpu_affine(ctx, name, h.msm(scalars[:m]))
        exp2, _ = oracle_affine(name, bases[:m], scalars[:m], inf=inf[: (m + 7) // 8].copy() if False else inf)
        assert (got2 == exp2).all()
        assert (h.download(0, 8) == bases[:8]).all()
    finally:
        h.free()


def test_reference_interface(ctx):
    """VariableBaseMSM::multi_scalar_mul mirror: size = min(len(bases), len(scalars))."""
    name = "bn254_g1"
    bases = cbind.bases_seq(name, 1, 100)
    scalars = random_scalars(80, curves.CURVES[name].fr.p, seed=3)
    res = ozl.ec.VariableBaseMSM.multi_scalar_mul(bases, scalars, curve=ozl.BN254_G1, ctx=ctx)
    got, inf = res.into_affine()
    exp, _ = oracle_affine(name, bases[:80], scalars)
    assert not inf and (got == exp).all()


def test_msm_submit_pipeline(ctx):
    """ozl_msm_submit: three back-to-back submissions with different scalars, one synchronize."""
    name = "bn254_g1"
    n = 5000
    r = curves.CURVES[name].fr.p
    bases = cbind.bases_seq(name, 1, n)
    h = ctx.upload_bases(ozl.BN254_G1, bases)
    try:
        scal = [random_scalars(n, r, seed=100 + i) for i in range(3)]
        outs = [np.zeros(12, dtype=np.uint64) for _ in range(3)]
        for s, o in zip(scal, outs):
            h.msm_submit(s.ctypes.data, n, o.ctypes.data)
        ctx.synchronize()
        for s, o in zip(scal, outs):
            got, _ = gpu_affine(ctx, name, o)
            assert (got == oracle_affine(name, bases, s)[0]).all()
    finally:
        h.free()


def test_jacobian_sum(ctx):
    name = "bls12_381_g1"
    ks = [5, 7, 0, 11]
    pts = np.stack([cbind.gen_mul(name, k) for k in ks])
    out = ctx.jacobian_sum(ozl.BLS12_381_G1, pts)
    got, _ = cbind.to_affine(name, out)
    exp, _ = cbind.to_affine(name, cbind.gen_mul(name, sum(ks)))
    assert (got == exp).all()


def test_error_paths(ctx):
    with pytest.raises(ozl.OzlError):
        ctx.upload_bases(99, np.zeros((1, 12), dtype=np.uint64))
    with pytest.raises(ozl.OzlError):
        ctx.upload_bases(ozl.BLS12_381_G1, np.zeros((1, 8), dtype=np.uint64))
    h = ctx.upload_bases(ozl.BN254_G1, cbind.bases_seq("bn254_g1", 1, 4))
    with pytest.raises(ozl.OzlError):
        h.msm(np.zeros((5, 4), dtype=np.uint64))      # more scalars than bases on a handle
    h.free()
    with pytest.raises(ozl.OzlError):
        h2 = ozl.Bases(ctx, 12345, ozl.BN254_G1, 4)
        h2.msm(np.zeros((1, 4), dtype=np.uint64))


@pytest.mark.parametrize("name", CURVES)
@pytest.mark.parametrize("levels", [1, 3, 5])
def test_msm_batch_affine_levels(ctx, name, levels):
    """Experimental batched-affine pair levels (msm_batch.cuh) give the same group element:
    directed scalars, and buckets built to hit every exceptional pair -- P + P (tangent),
    P + (-P) (cancellation), a cancelled pair meeting a live one on the next level, and
    buckets that cancel completely."""
    n = 1 << 11 if name.endswith("g1") else 1 << 9
    bases = cbind.bases_seq(name, 11, n)
    c = curves.CURVES[name]
    neg = lambda row: np.array(c.affine_to_mont_limbs(c.neg(c.affine_from_mont_limbs(list(row)))), dtype=np.uint64)
    scalars = directed_scalars(name, n, seed=levels)
    # same scalar -> same bucket in every window; sorted order inside a bucket is not fixed, so use
    # groups where ANY pairing hits an exceptional case
    k = 40
    for j in range(1, 4):
        bases[k + j] = bases[k]              # four copies of P: tangent pairs at two levels
        scalars[k + j] = scalars[k]
    k = 60
    bases[k + 1] = neg(bases[k])             # P, -P, P, -P: cancels whatever the pairing
    bases[k + 2] = bases[k]
    bases[k + 3] = neg(bases[k])
    for j in range(1, 4):
        scalars[k + j] = scalars[k]
    k = 80
    bases[k + 1] = neg(bases[k])             # P, -P, Q: identity meets a live point
    scalars[k + 1] = scalars[k]
    scalars[k + 2] = scalars[k]
    exp, exp_inf = oracle_affine(name, bases, scalars)
    h = ctx.upload_bases(ozl.CURVE_IDS[name], bases)
    try:
        ctx.set_window_bits(6)               # few buckets -> long lists, several live levels
        ctx.set_batch_affine(levels)
        got, got_inf = gpu_affine(ctx, name, h.msm(scalars))
    finally:
        ctx.set_batch_affine(-1)
        ctx.set_window_bits(0)
        h.free()
    assert got_inf == exp_inf
    assert (got == exp).all()


def test_msm_batch_affine_all_equal_and_tiny(ctx):
    """One bucket holding every point (all scalars equal) and sizes around the pairing edge cases."""
    name = "bn254_g1"
    r = curves.CURVES[name].fr.p
    try:
        ctx.set_batch_affine(4)
        for n in (1, 2, 3, 5, 33, 1000):
            bases = cbind.bases_seq(name, 5, n)
            scalars = np.repeat(ints_to_array([0x1234567 % r]), n, axis=0)
            exp, exp_inf = oracle_affine(name, bases, scalars)
            h = ctx.upload_bases(ozl.BN254_G1, bases)
            try:
                got, got_inf = gpu_affine(ctx, name, h.msm(scalars))
            finally:
                h.free()
            assert got_inf == exp_inf and (got == exp).all(), n
    finally:
        ctx.set_batch_affine(-1)


@pytest.mark.parametrize("name", ["bls12_381_g1", "bn254_g2"])
def test_msm_precompute_one_copy_per_window(ctx, name):
    """A factor above the number of windows (what bench.py and the Groth16 bench pass: 32) means one
    copy per window: a single bucket set, no Horner tail, same group element."""
    n = 2000 if name.endswith("g1") else 500
    bases = cbind.bases_seq(name, 9, n)
    scalars = directed_scalars(name, n, seed=77)
    exp, _ = oracle_affine(name, bases, scalars)
    h = ctx.upload_bases(ozl.CURVE_IDS[name], bases).precompute(32)
    try:
        info = h.info(n)
        assert info["bucket_sets"] == 1 and info["factor"] == info["windows"]
        got, _ = gpu_affine(ctx, name, h.msm(scalars))
        assert (got == exp).all()
    finally:
        h.free()


def _known_dlog_affine(name, scalars, start, n):
    k = cbind.dot_mod_r("bls12_381_fr", scalars, np.arange(start, start + n, dtype=np.uint64))
    return cbind.to_affine(name, cbind.gen_mul(name, k))[0]


@pytest.mark.parametrize("precompute", [1, 32])
def test_msm_host_batches_pinned_pageable_device_agree(ctx, precompute):
    """`ozl_msm` with page-locked scalars (3 batches), with pageable scalars (7 batches) and
    `ozl_msm_device_async` with resident scalars (one batch) return the same point at 2^22, with and
    without one shifted copy of the bases per window; a short prefix (n < uploaded count) too."""
    import torch
    name = "bls12_381_g1"
    n = 1 << 22
    r = curves.CURVES[name].fr.p
    scalars = random_scalars(n, r, seed=77)
    h = ctx.generate_bases(ozl.BLS12_381_G1, 5, n)
    try:
        if precompute > 1:
            h.precompute(precompute)
        exp = _known_dlog_affine(name, scalars, 5, n)
        got_pageable, _ = gpu_affine(ctx, name, h.msm(scalars))
        pinned = torch.from_numpy(scalars.view(np.int64)).pin_memory()
        out = np.zeros(18, dtype=np.uint64)
        h.msm_host_ptr(pinned.data_ptr(), n, out)
        got_pinned, _ = gpu_affine(ctx, name, out)
        d_s = pinned.cuda()
        d_out = torch.zeros(18, dtype=torch.int64, device="cuda")
        torch.cuda.synchronize()
        h.msm_device(d_s.data_ptr(), n, d_out.data_ptr())
        ctx.synchronize()
        got_dev, _ = gpu_affine(ctx, name, d_out.cpu().numpy().view(np.uint64))
        assert (got_pageable == exp).all() and (got_pinned == exp).all() and (got_dev == exp).all()
        m = n - 12345                               # ragged prefix through the batched path
        got_m, _ = gpu_affine(ctx, name, h.msm(scalars[:m]))
        assert (got_m == _known_dlog_affine(name, scalars[:m], 5, m)).all()
    finally:
        h.free()


def test_msm_host_batches_skewed(ctx):
    """Batched accumulation with heavy buckets in every batch: a witness-like scalar vector (mostly
    small values, many equal) at 2^22 through the pageable 7-batch path."""
    name = "bls12_381_g1"
    n = 1 << 22
    r = curves.CURVES[name].fr.p
    rng = np.random.default_rng(5)
    scalars = np.zeros((n, 4), dtype=np.uint64)
    scalars[:, 0] = rng.integers(0, 4, size=n, dtype=np.uint64)            # three quarters of the points in three buckets
    full = random_scalars(n // 8, r, seed=6)
    scalars[::8] = full
    h = ctx.generate_bases(ozl.BLS12_381_G1, 1, n)
    try:
        got, _ = gpu_affine(ctx, name, h.msm(scalars))
        assert (got == _known_dlog_affine(name, scalars, 1, n)).all()
    finally:
        h.free()


def test_msm_rejects_non_canonical_scalars(ctx):
    """ark's multi_scalar_mul takes any BigInteger256; this library's window plan covers canonical
    scalars only and must say so (OZL_ERR_ARG) instead of returning a wrong point."""
    name = "bn254_g1"
    n = 64
    bases = cbind.bases_seq(name, 1, n)
    r = curves.CURVES[name].fr.p
    scalars = random_scalars(n, r, seed=9)
    h = ctx.upload_bases(ozl.BN254_G1, bases)
    try:
        h.msm(scalars)                               # canonical: fine
        bad = scalars.copy()
        bad[17, 3] |= np.uint64(1 << 63)             # bit 255 set: not reduced modulo r
        with pytest.raises(ozl.OzlError) as ei:
            h.msm(bad)
        assert ei.value.status == 1
        got, _ = gpu_affine(ctx, name, h.msm(scalars))      # the context stays usable
        assert (got == oracle_affine(name, bases, scalars)[0]).all()
    finally:
        h.free()


@pytest.mark.parametrize("factor", [1, 4])
def test_msm_zero_zero_base_is_identity(ctx, factor):
    """(0, 0) encodes the point at infinity even without an inf_mask, with and without precomputed copies."""
    name = "bls12_381_g1"
    n = 500
    bases = cbind.bases_seq(name, 1, n)
    bases[3] = 0
    bases[499] = 0
    scalars = directed_scalars(name, n, seed=21)
    inf = np.zeros((n + 7) // 8, dtype=np.uint8)
    inf[3 >> 3] |= 1 << (3 & 7)
    inf[499 >> 3] |= 1 << (499 & 7)
    exp, _ = oracle_affine(name, bases, scalars, inf=inf)
    h = ctx.upload_bases(ozl.BLS12_381_G1, bases)         # no mask passed
    try:
        if factor > 1:
            h.precompute(factor)
        got, _ = gpu_affine(ctx, name, h.msm(scalars))
        assert (got == exp).all()
    finally:
        h.free()
