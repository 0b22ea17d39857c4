"""Host-side constants and helpers of the ``EvaluationDomain`` mirror (ark_poly 0.3.0 trait methods that
need no device): group_gen, size_inv, element, vanishing polynomial, Lagrange coefficients --
against the oracle's domain and against each other.  CPU only; no Context is created."""
import random

import pytest

from openzl_b200 import _lib
from openzl_b200.poly import Radix2EvaluationDomain
from oracle import fields
from oracle.ntt import Radix2Domain as Domain

CASES = [(_lib.BN254_FR, fields.BN254_FR), (_lib.BLS12_381_FR, fields.BLS12_381_FR)]


def _dom(fid, n):
    return Radix2EvaluationDomain(fid, n, n.bit_length() - 1, ctx=None)


@pytest.mark.parametrize("fid,f", CASES)
@pytest.mark.parametrize("log_n", [0, 1, 5, 12, 20])
def test_domain_constants_match_the_oracle(fid, f, log_n):
    n = 1 << log_n
    d, o = _dom(fid, n), Domain(f, n)
    p = f.p
    assert d.modulus == p and d.size() == o.size
    assert d.group_gen == o.group_gen and d.group_gen_inv == o.group_gen_inv
    assert d.size_inv == o.size_inv and d.generator_inv == o.generator_inv
    assert pow(d.group_gen, n, p) == 1 and (n == 1 or pow(d.group_gen, n // 2, p) == p - 1)   # primitive n-th root
    for i in (0, 1, n - 1, 12345 % n):
        assert d.element(i) == o.element(i)
    tau = random.Random(log_n).randrange(p)
    assert d.evaluate_vanishing_polynomial(tau) == o.evaluate_vanishing_polynomial(tau)


@pytest.mark.parametrize("fid,f", CASES)
def test_lagrange_coefficients(fid, f):
    """sum_i L_i(tau) x_i equals the interpolating polynomial's value at tau (coefficients by the
    oracle's ifft), and L(omega^k) is the k-th indicator."""
    p, n = f.p, 64
    d, o = _dom(fid, n), Domain(f, n)
    rnd = random.Random(7)
    evals = [rnd.randrange(p) for _ in range(n)]
    coeffs = o.ifft(evals)
    tau = rnd.randrange(p)
    L = d.evaluate_all_lagrange_coefficients(tau)
    horner = 0
    for c in reversed(coeffs):
        horner = (horner * tau + c) % p
    assert sum(l * x for l, x in zip(L, evals)) % p == horner
    assert sum(L) % p == 1                                   # partition of unity
    k = 9
    ind = d.evaluate_all_lagrange_coefficients(d.element(k))
    assert ind == [1 if i == k else 0 for i in range(n)]
    assert list(d.elements())[:4] == [o.element(i) for i in range(4)]
