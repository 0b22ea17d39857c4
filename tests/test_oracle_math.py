"""Mathematical cross-checks that make the oracle self-validating where the reference pins
nothing: ark-restated Pippenger == naive sum == known-dlog
identity; radix-2 NTT == O(n^2) DFT; inverse transforms are inverses; C++ oracle == Python oracle.
"""
import random

import numpy as np
import pytest

from oracle import cbind, curves, fields
from oracle.msm import ark_window_bits, msm_ark, msm_known_dlog, msm_naive
from oracle.ntt import Radix2Domain
from tests.util import ints_to_array, limbs_to_int, random_scalars, scalars_to_ints

ALL = list(curves.CURVES)


def test_ark_window_rule():
    # table verified in SURVEY.md appendix
    assert [ark_window_bits(1 << k) for k in (12, 20, 22, 24, 26, 28)] == [10, 15, 17, 18, 19, 21]
    assert ark_window_bits(31) == 3 and ark_window_bits(32) == 5
    assert [cbind.window_bits(1 << k) for k in (12, 20, 26)] == [10, 15, 19]


@pytest.mark.parametrize("name", ALL)
def test_group_law_formulas(name):
    c = curves.CURVES[name]
    rnd = random.Random(1)
    P = c.mul_affine(c.gen, rnd.randrange(1, c.fr.p))
    Q = c.mul_affine(c.gen, rnd.randrange(1, c.fr.p))
    assert c.is_on_curve(P) and c.is_on_curve(Q)
    assert c.to_affine(c.add_mixed(c.to_jac(P), Q)) == c.add_affine(P, Q)
    assert c.to_affine(c.add_jac(c.dbl_jac(c.to_jac(P)), c.to_jac(Q))) == c.add_affine(c.add_affine(P, P), Q)
    assert c.to_affine(c.add_mixed(c.to_jac(P), P)) == c.add_affine(P, P)          # doubling guard
    assert c.to_affine(c.add_mixed(c.to_jac(P), c.neg(P))) is None                 # inverse guard
    assert c.to_affine(c.add_mixed(c.identity_jac(), P)) == P


def test_public_known_answer_doublings():
    """Widely published doubling vectors (NOT from /root/reference, which holds no curve vectors):
    2*G1 on BLS12-381 (zkcrypto/bls12_381 and the IETF pairing-friendly-curves draft) and
    (1,2)+(1,2) on alt_bn128 (EIP-196 ecAdd test vector).  They pin the oracle's group law to
    something outside this repository; the C++ oracle and the device are then compared with it."""
    c = curves.BLS12_381_G1
    assert c.mul_affine(c.gen, 2) == (
        0x0572CBEA904D67468808C8EB50A9450C9721DB309128012543902D0AC358A62AE28F75BB8F1C7C42C39A8C5529BF0F4E,
        0x166A9D8CABC673A322FDA673779D8E3822BA3ECB8670E461F73BB9021D5FD76A4C56D9D4CD16BD1BBA86881979749D28)
    b = curves.BN254_G1
    assert b.add_affine(b.gen, b.gen) == (
        1368015179489954701390400359078579693043519447331113978918064868415326638035,
        9918110051302171585080402603319702774565515993150576347155970296011118125764)
    for name, cv in (("bls12_381_g1", c), ("bn254_g1", b)):
        aff, inf = cbind.to_affine(name, cbind.gen_mul(name, 2))
        assert not inf and cv.affine_from_mont_limbs(list(aff)) == cv.mul_affine(cv.gen, 2)


@pytest.mark.parametrize("name", ALL)
def test_msm_ark_equals_naive_and_dlog(name):
    c = curves.CURVES[name]
    rnd = random.Random(2)
    n = 40
    d = [rnd.randrange(1, 1 << 40) for _ in range(n)]
    bases = [c.mul_affine(c.gen, x) for x in d]
    s = [rnd.randrange(c.fr.p) for _ in range(n)]
    s[0], s[1], s[2] = 0, 1, c.fr.p - 1
    bases[3], d[3] = None, 0
    bases[5], d[5], s[5] = bases[4], d[4], s[4]
    bases[7], d[7], s[7] = c.neg(bases[6]), c.fr.p - d[6], s[6]
    a = c.to_affine(msm_ark(c, bases, s))
    assert a == c.to_affine(msm_naive(c, bases, s))
    assert a == c.to_affine(msm_known_dlog(c, d, s))


@pytest.mark.parametrize("name", ALL)
def test_c_oracle_msm_equals_python(name):
    c = curves.CURVES[name]
    n = 64
    bases = cbind.bases_seq(name, 1, n)
    assert list(bases[6]) == c.affine_to_mont_limbs(c.mul_affine(c.gen, 7))
    s = random_scalars(n, c.fr.p, seed=5)
    s[0] = 0
    s[1] = ints_to_array([1])[0]
    s[2] = ints_to_array([c.fr.p - 1])[0]
    inf = np.zeros(8, dtype=np.uint8)
    inf[1] = 1  # point 8 at infinity
    py_bases = [c.mul_affine(c.gen, i + 1) for i in range(n)]
    py_bases[8] = None
    exp = c.to_affine(msm_ark(c, py_bases, scalars_to_ints(s)))
    for threads in (1, 4):
        aff, is_inf = cbind.to_affine(name, cbind.msm(name, bases, s, inf=inf, threads=threads))
        assert not is_inf and c.affine_from_mont_limbs(list(aff)) == exp
    k = cbind.dot_mod_r(c.fr.name, s, np.arange(1, n + 1, dtype=np.uint64))
    assert k == sum(v * (i + 1) for i, v in enumerate(scalars_to_ints(s))) % c.fr.p


def test_c_oracle_known_dlog_2_14():
    name = "bls12_381_g1"
    n = 1 << 14
    bases = cbind.bases_seq(name, 1, n)
    s = random_scalars(n, curves.BLS12_381_G1.fr.p, seed=11)
    got, _ = cbind.to_affine(name, cbind.msm(name, bases, s, threads=8))
    k = cbind.dot_mod_r("bls12_381_fr", s, np.arange(1, n + 1, dtype=np.uint64))
    exp, _ = cbind.to_affine(name, cbind.gen_mul(name, k))
    assert (got == exp).all()


@pytest.mark.parametrize("fname", ["bls12_381_fq", "bls12_381_fr", "bn254_fq", "bn254_fr"])
def test_c_oracle_field_mul(fname):
    f = fields.FIELDS[fname]
    rnd = random.Random(3)
    vals = [0, 1, f.p - 1, f.p - 2, (1 << (64 * f.limbs64 - 1)) % f.p, f.R, f.R2] + [rnd.randrange(f.p) for _ in range(200)]
    for a in vals[:12]:
        for b in vals:
            out = cbind.fp_mul(fname, np.array(f.to_limbs(a), dtype=np.uint64), np.array(f.to_limbs(b), dtype=np.uint64))
            assert limbs_to_int(out) == f.mont_mul(a, b)


@pytest.mark.parametrize("fname", ["bn254_fr", "bls12_381_fr"])
@pytest.mark.parametrize("log_n", [0, 1, 3, 6])
def test_ntt_equals_dft(fname, log_n):
    f = fields.FIELDS[fname]
    rnd = random.Random(log_n)
    n = 1 << log_n
    d = Radix2Domain(f, n)
    x = [rnd.randrange(f.p) for _ in range(n)]
    assert d.fft(x) == d.dft_naive(x)
    assert d.ifft(x) == d.dft_naive(x, inverse=True)
    assert d.ifft(d.fft(x)) == x
    assert d.coset_ifft(d.coset_fft(x)) == x
    X = np.array([f.to_limbs(f.to_mont(v)) for v in x], dtype=np.uint64).reshape(n, 4)
    for inv, cos, fn in [(False, False, d.fft), (True, False, d.ifft), (False, True, d.coset_fft), (True, True, d.coset_ifft)]:
        y = cbind.ntt(fname, X, inverse=inv, coset=cos)
        assert [f.from_mont(limbs_to_int(r)) for r in y] == fn(x)


def test_domain_sizing_like_ark():
    f = fields.BN254_FR
    assert Radix2Domain(f, 1000).size == 1024 and Radix2Domain(f, 1024).size == 1024 and Radix2Domain(f, 1025).size == 2048
    with pytest.raises(ValueError):
        Radix2Domain(f, (1 << 28) + 1)
    # coset evaluation really is evaluation on g*H: spot-check one point
    d = Radix2Domain(f, 8)
    coeffs = [3, 1, 4, 1, 5, 9, 2, 6]
    ev = d.coset_fft(coeffs)
    pt = (f.generator * d.element(3)) % f.p
    assert ev[3] == sum(c * pow(pt, i, f.p) for i, c in enumerate(coeffs)) % f.p
    # vanishing polynomial on the coset is the constant g^n - 1
    assert d.evaluate_vanishing_polynomial(f.generator) == (pow(f.generator, 8, f.p) - 1) % f.p
