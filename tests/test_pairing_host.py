"""CPU tests of the product's host-side pairing / `Groth16.verify` (openzl_b200/pairing.py), the mirror
of the plugin's `Pairing` trait and `ProofSystem::verify`
(/root/reference/plugins/arkworks/src/pairing.rs:9-90, groth16.rs:459-466).  Like the reference's own
pairing test (pairing.rs:104-129) the pin is bilinearity; on top of that the product's ate pairing and
the oracle's Tate pairing -- two independent implementations -- must take the same verification
decisions on the same Groth16 instances."""
import random

import numpy as np
import pytest

from openzl_b200 import pairing as pr
from openzl_b200.circuits import PoseidonChain
from openzl_b200.groth16 import Groth16, Proof, Trapdoor, VerifyingData, _uniform_scalar
from oracle import curves, fields
from oracle import groth16 as og
from oracle import pairing as opair


@pytest.mark.parametrize("name", ["bn254", "bls12_381"])
def test_bilinearity_and_non_degeneracy(name):
    E = pr.ENGINES[name]
    rnd = random.Random(len(name))
    a, b = rnd.randrange(1, E.r), rnd.randrange(1, E.r)
    e = pr.pairing(E, E.g1, E.g2)
    assert e != pr._f12_one() and pr._f12_pow(e, E.r, E) == pr._f12_one()
    assert pr.pairing(E, pr.g1_mul(E, E.g1, a), pr.g2_mul(E, E.g2, b)) == pr._f12_pow(e, a * b % E.r, E)
    assert pr.pairing(E, None, E.g2) == pr._f12_one() and pr.pairing(E, E.g1, None) == pr._f12_one()
    # the equation form used by verify
    aG, bH = pr.g1_mul(E, E.g1, a), pr.g2_mul(E, E.g2, b)
    assert pr.product_of_pairings_is_one(E, [(aG, bH), (pr.g1_neg(E, pr.g1_mul(E, E.g1, a * b)), E.g2)])
    assert not pr.product_of_pairings_is_one(E, [(aG, bH), (pr.g1_neg(E, pr.g1_mul(E, E.g1, a * b + 1)), E.g2)])


@pytest.mark.parametrize("name", ["bn254", "bls12_381"])
def test_valid_pairing_ratio_like_the_reference(name):
    """The reference's own pairing test, restated (`assert_valid_pairing_ratio`, pairing.rs:104-129):
    (g1, scalar * g2) and (scalar * g1, g2) evaluate to the same element -- through the mirrored
    `PairingEngineExt::{same, has_same, same_ratio}` -- for random points of both groups."""
    E = pr.ENGINES[name]
    rnd = random.Random(7 + len(name))
    for _ in range(2):
        g1 = pr.g1_mul(E, E.g1, rnd.randrange(1, E.r))          # a random G1 / G2 element, like rng.gen()
        g2 = pr.g2_mul(E, E.g2, rnd.randrange(1, E.r))
        scalar = rnd.randrange(1, E.r)
        lhs, rhs = (g1, pr.g2_mul(E, g2, scalar)), (pr.g1_mul(E, g1, scalar), g2)
        assert pr.same(E, lhs, rhs) == (lhs, rhs)
        assert pr.same(E, lhs, (pr.g1_mul(E, g1, scalar + 1), g2)) is None
        # same_ratio((g1, s g1), (g2, s g2))
        assert pr.same_ratio(E, (g1, pr.g1_mul(E, g1, scalar)), (g2, pr.g2_mul(E, g2, scalar)))
        assert not pr.same_ratio(E, (g1, pr.g1_mul(E, g1, scalar)), (g2, pr.g2_mul(E, g2, scalar + 1)))


def test_generators_match_the_oracle_curves():
    for name, E in pr.ENGINES.items():
        g1, g2 = curves.CURVES[name + "_g1"], curves.CURVES[name + "_g2"]
        assert tuple(g1.gen) == E.g1
        assert pr.g1_on_curve(E, E.g1) and pr.g2_on_curve(E, E.g2)
        k = 0xDEADBEEF12345
        assert g1.mul_affine(g1.gen, k) == pr.g1_mul(E, E.g1, k)
        assert pr.g1_mul(E, E.g1, E.r) is None and pr.g2_mul(E, E.g2, E.r) is None


def _limbs(c, pt):
    return np.array(c.affine_to_mont_limbs(pt), dtype=np.uint64)


@pytest.mark.parametrize("name,fname", [("bn254", "bn254_fr"), ("bls12_381", "bls12_381_fr")])
def test_groth16_verify_agrees_with_oracle(name, fname):
    """A proof built from its discrete logs by the CPU group law verifies under BOTH pairing
    implementations; tampered proofs / inputs are rejected by both."""
    from openzl_b200.circuits import PoseidonParams
    f = fields.FIELDS[fname]
    p = f.p
    ch = PoseidonChain(1) if name == "bn254" else PoseidonChain(1, PoseidonParams.generate(modulus=p))
    r1 = ch.r1cs()
    z = ch.assignment(3, 4)
    rnd = random.Random(11)
    td = Trapdoor(*[rnd.randrange(2, p) for _ in range(5)])
    A, B, C = og.prove_exponents(fname, r1, z, td, rnd.randrange(p), rnd.randrange(p))
    g1, g2 = curves.CURVES[name + "_g1"], curves.CURVES[name + "_g2"]
    a, b, c, _ = og.qap_at_tau(fname, r1, td.tau)
    ginv = f.inv(td.gamma)
    ic = [(td.beta * a[j] + td.alpha * b[j] + c[j]) % p * ginv % p for j in range(r1.n_instance)]
    vk = VerifyingData(name, td, ic, _limbs(g1, g1.mul_affine(g1.gen, td.alpha)), _limbs(g2, g2.mul_affine(g2.gen, td.beta)),
                       _limbs(g2, g2.mul_affine(g2.gen, td.gamma)), _limbs(g2, g2.mul_affine(g2.gen, td.delta)),
                       np.stack([_limbs(g1, g1.mul_affine(g1.gen, k)) for k in ic]))
    pts = (g1.mul_affine(g1.gen, A), g2.mul_affine(g2.gen, B), g1.mul_affine(g1.gen, C))
    proof = Proof(_limbs(g1, pts[0]), _limbs(g2, pts[1]), _limbs(g1, pts[2]))
    assert Groth16.verify(vk, [z[1]], proof)
    assert not Groth16.verify(vk, [(z[1] + 1) % p], proof)
    assert not Groth16.verify(vk, [], proof)                       # wrong number of public inputs
    bad = Proof(proof.a, proof.b, _limbs(g1, g1.mul_affine(g1.gen, (C + 1) % p)))
    assert not Groth16.verify(vk, [z[1]], bad)
    off_curve = Proof(proof.a.copy(), proof.b, proof.c)
    off_curve.a[0] ^= np.uint64(1)
    assert not Groth16.verify(vk, [z[1]], off_curve)
    # the oracle's (Tate) pairing takes the same decisions
    ic_pts = [g1.mul_affine(g1.gen, k) for k in ic]
    o = lambda x, pr_: opair.groth16_verify(name, g1.mul_affine(g1.gen, td.alpha), g2.mul_affine(g2.gen, td.beta),
                                            g2.mul_affine(g2.gen, td.gamma), g2.mul_affine(g2.gen, td.delta), ic_pts, x, pr_)
    assert o([z[1]], pts) and not o([(z[1] + 1) % p], pts)
    # vk wire format round trip (ark compressed VerifyingKey) keeps verifying
    from openzl_b200 import serialize as ser
    vk2 = VerifyingData.from_serializable(name, ser.vk_from_bytes(name, vk.to_bytes())[0])
    assert Groth16.verify(vk2, [z[1]], proof)


def test_blinding_scalars_are_uniform_and_need_a_csprng():
    p = fields.BN254_FR.p
    vals = [_uniform_scalar(p) for _ in range(64)]
    assert all(0 <= v < p for v in vals) and len(set(vals)) == 64
    assert max(vals).bit_length() >= 250                     # full-width draws, not 62 bits of entropy
    assert 0 <= _uniform_scalar(p, random.SystemRandom()) < p
    with pytest.raises(TypeError):
        _uniform_scalar(p, random.Random(1))
    with pytest.raises(TypeError):
        _uniform_scalar(p, np.random.default_rng(1))
