"""The oracle's pairing (test infrastructure): bilinear, non-degenerate, and the Groth16 verification
equation holds with REAL pairings for the proof a correct prover outputs -- the check the reference does
with `verify_with_processed_vk` (/root/reference/plugins/arkworks/src/groth16.rs:460-466) and its own
bilinearity tests (/root/reference/plugins/arkworks/src/pairing.rs:104-129).  CPU only."""
import random

import pytest

from openzl_b200.circuits import PoseidonChain, PoseidonParams
from openzl_b200.groth16 import Trapdoor
from oracle import curves, fields, pairing as opair
from oracle import groth16 as og


@pytest.mark.parametrize("name", ["bn254", "bls12_381"])
def test_bilinear_and_non_degenerate(name):
    P = opair.PAIRINGS[name]
    F = opair.Fq12(P)
    g1, g2 = P.g1, P.g2
    r = g1.fr.p
    rnd = random.Random(11)
    a, b = rnd.randrange(2, 1 << 40), rnd.randrange(2, 1 << 40)
    e = opair.pairing(name, g1.gen, g2.gen)
    assert e != F.one                                             # non-degenerate
    assert F.pow(e, r) == F.one                                   # lands in mu_r
    e_ab = opair.pairing(name, g1.mul_affine(g1.gen, a), g2.mul_affine(g2.gen, b))
    assert e_ab == F.pow(e, (a * b) % r)                          # e(aP, bQ) = e(P, Q)^(ab), pairing.rs:104-129
    assert opair.pairing(name, g1.mul_affine(g1.gen, a), g2.gen) == opair.pairing(name, g1.gen, g2.mul_affine(g2.gen, a))
    # product form used by the verifier
    assert opair.product_of_pairings_is_one(name, [(g1.neg(g1.mul_affine(g1.gen, a)), g2.gen), (g1.gen, g2.mul_affine(g2.gen, a))])
    assert not opair.product_of_pairings_is_one(name, [(g1.gen, g2.gen), (g1.gen, g2.mul_affine(g2.gen, a))])
    assert opair.pairing(name, None, g2.gen) == F.one


@pytest.mark.parametrize("name,fname", [("bn254", "bn254_fr"), ("bls12_381", "bls12_381_fr")])
def test_groth16_equation_with_real_pairings(name, fname):
    """Points [A']G1, [B']G2, [C']G1 of the oracle prover satisfy e(A,B) = e(alpha,beta) e(IC,gamma) e(C,delta);
    a wrong public input or a perturbed proof does not.  (The GPU tests show the device's proof points
    are exactly these points.)"""
    f = fields.FIELDS[fname]
    p = f.p
    ch = PoseidonChain(1) if name == "bn254" else PoseidonChain(1, PoseidonParams.generate(modulus=p))
    r1 = ch.r1cs()
    z = ch.assignment(5, 6)
    rnd = random.Random(3)
    td = Trapdoor(*[rnd.randrange(2, p) for _ in range(5)])
    r, s = rnd.randrange(p), rnd.randrange(p)
    A, B, C = og.prove_exponents(fname, r1, z, td, r, s)
    P = opair.PAIRINGS[name]
    g1, g2 = P.g1, P.g2
    a, b, c, _ = og.qap_at_tau(fname, r1, td.tau)
    ginv = f.inv(td.gamma)
    ic = [g1.mul_affine(g1.gen, (td.beta * a[j] + td.alpha * b[j] + c[j]) % p * ginv % p) for j in range(r1.n_instance)]
    vk = dict(alpha_g1=g1.mul_affine(g1.gen, td.alpha), beta_g2=g2.mul_affine(g2.gen, td.beta),
              gamma_g2=g2.mul_affine(g2.gen, td.gamma), delta_g2=g2.mul_affine(g2.gen, td.delta), gamma_abc_g1=ic)
    proof = (g1.mul_affine(g1.gen, A), g2.mul_affine(g2.gen, B), g1.mul_affine(g1.gen, C))
    assert opair.groth16_verify(name, public_inputs=[z[1]], proof=proof, **vk)
    assert not opair.groth16_verify(name, public_inputs=[(z[1] + 1) % p], proof=proof, **vk)
    bad = (proof[0], proof[1], g1.mul_affine(g1.gen, (C + 1) % p))
    assert not opair.groth16_verify(name, public_inputs=[z[1]], proof=bad, **vk)


def test_verify_helper_on_abi_layout():
    """tests/util.verify_proof_with_pairings (what the GPU Groth16 test calls on the device's output),
    fed here with the oracle prover's points in the same limb layout."""
    import numpy as np
    from tests.util import verify_proof_with_pairings
    p = fields.BN254_FR.p
    ch = PoseidonChain(1)
    r1 = ch.r1cs()
    z = ch.assignment(8, 9)
    rnd = random.Random(21)
    td = Trapdoor(*[rnd.randrange(2, p) for _ in range(5)])
    A, B, C = og.prove_exponents("bn254_fr", r1, z, td, rnd.randrange(p), rnd.randrange(p))
    g1, g2 = curves.BN254_G1, curves.BN254_G2
    limbs = lambda g, k: np.array(g.affine_to_mont_limbs(g.mul_affine(g.gen, k)), dtype=np.uint64)
    pa, pb, pc = limbs(g1, A), limbs(g2, B), limbs(g1, C)
    assert verify_proof_with_pairings("bn254", "bn254_fr", r1, td, [z[1]], pa, pb, pc)
    assert not verify_proof_with_pairings("bn254", "bn254_fr", r1, td, [z[1]], pa, pb, limbs(g1, C + 1))
    assert not verify_proof_with_pairings("bn254", "bn254_fr", r1, td, [z[1]], pa, pb, np.ones_like(pc))   # not on the curve
