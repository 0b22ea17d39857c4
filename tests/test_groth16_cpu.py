"""CPU-side checks of the Groth16 row: the circuit generator (host logic of the product) and the
oracle's restatement of ark_groth16::create_proof, tied together by the verification equation in
the exponent.  No GPU, no C-ABI compute calls."""
import random

import numpy as np
import pytest

from openzl_b200.circuits import PoseidonChain, PoseidonParams
from openzl_b200.groth16 import Trapdoor, ints_to_limbs, limbs_to_ints
from oracle import fields, poseidon as opos
from oracle import groth16 as og

P = fields.BN254_FR.p


def test_params_match_oracle_and_reference_kat():
    # product-side Poseidon constants == oracle's (which is pinned to the reference's golden vectors)
    pr = fields.BLS12_381_FR
    pp = PoseidonParams.generate(modulus=pr.p)
    assert pp.round_keys == opos.generate_round_constants(pr, 3, 8, 55)
    assert pp.mds == opos.generate_mds(pr, 3)
    assert pp.permute([3, 1, 2]) == opos.permute(pr, [3, 1, 2], pp.round_keys, pp.mds, 8, 55)
    assert pp.permute([3, 1, 2])[0] == 1808609226548932412441401219270714120272118151392880709881321306315053574086
    # BN254 instantiation the plugin configures: arity 2 -> width 3, 8 full + 55 partial rounds
    pb = PoseidonParams.generate()
    assert (pb.width, pb.full_rounds, pb.partial_rounds, pb.domain_tag) == (3, 8, 55, 3)
    assert len(pb.round_keys) == 189


@pytest.mark.parametrize("links", [1, 2])
def test_chain_r1cs_shape_and_satisfaction(links):
    ch = PoseidonChain(links)
    r1 = ch.r1cs()
    assert ch.constraints_per_link == 8 * 9 + 55 * 5 + 1 == 348
    assert r1.n_constraints == 348 * links and r1.n_vars == 4 + 348 * links and r1.n_instance == 2
    z = ch.assignment(17, 23)
    assert z[0] == 1 and z[1] == ch.digest(17, 23)
    assert r1.is_satisfied(z)
    bad = list(z)
    bad[1] = (bad[1] + 1) % P
    assert not r1.is_satisfied(bad)
    bad = list(z)
    bad[2] = (bad[2] + 1) % P          # wrong preimage
    assert not r1.is_satisfied(bad)
    # CSR sanity: indices in range, every constraint defines exactly one C entry
    for M in (r1.A, r1.B, r1.C):
        assert M.row_ptr[0] == 0 and (np.diff(M.row_ptr.astype(np.int64)) >= 1).all()
        assert M.col_idx.max() < r1.n_vars and M.coef_idx.max() < len(r1.coef_table)
    assert (np.diff(r1.C.row_ptr.astype(np.int64)) == 1).all()
    # transpose used by the setup is a true transpose
    At = r1.A.transpose(r1.n_vars)
    dense = {}
    for r in range(r1.A.n_rows):
        for k in range(int(r1.A.row_ptr[r]), int(r1.A.row_ptr[r + 1])):
            dense[(r, int(r1.A.col_idx[k]))] = int(r1.A.coef_idx[k])
    cnt = 0
    for c in range(At.n_rows):
        for k in range(int(At.row_ptr[c]), int(At.row_ptr[c + 1])):
            assert dense[(int(At.col_idx[k]), c)] == int(At.coef_idx[k])
            cnt += 1
    assert cnt == len(dense)


def test_oracle_proof_verifies_in_the_exponent():
    ch = PoseidonChain(1)
    r1 = ch.r1cs()
    z = ch.assignment(5, 6)
    rnd = random.Random(3)
    td = Trapdoor(*[rnd.randrange(2, P) for _ in range(5)])
    h = og.witness_map("bn254_fr", r1, z)
    n = len(h)
    assert n == 512 and h[n - 1] == 0
    # h really is the quotient: A(x)B(x) - C(x) = h(x) Z(x) at a random point
    a, b, c, _ = og.qap_at_tau("bn254_fr", r1, td.tau)
    az = sum(x * y for x, y in zip(z, a)) % P
    bz = sum(x * y for x, y in zip(z, b)) % P
    cz = sum(x * y for x, y in zip(z, c)) % P
    ht = sum(hi * pow(td.tau, i, P) for i, hi in enumerate(h)) % P
    zt = (pow(td.tau, n, P) - 1) % P
    # ark folds the instance rows into a(x); they multiply b(x) = 0 there, so the identity is unchanged
    assert (az * bz - cz - ht * zt) % P == 0
    for r, s in ((0, 0), (rnd.randrange(P), rnd.randrange(P))):
        A, B, C = og.prove_exponents("bn254_fr", r1, z, td, r, s, h=h)
        assert og.verify_exponents("bn254_fr", r1, td, [z[1]], A, B, C)
        assert not og.verify_exponents("bn254_fr", r1, td, [(z[1] + 1) % P], A, B, C)
        assert not og.verify_exponents("bn254_fr", r1, td, [z[1]], A, B, (C + 1) % P)


def test_limb_helpers_round_trip():
    rnd = random.Random(1)
    vals = [0, 1, P - 1] + [rnd.randrange(P) for _ in range(50)]
    assert limbs_to_ints(ints_to_limbs(vals)) == vals
    assert limbs_to_ints(ints_to_limbs(vals, P, mont=True), P, mont=True) == vals
    f = fields.BN254_FR
    assert [int(x) for x in ints_to_limbs([5], P, mont=True)[0]] == f.to_limbs(f.to_mont(5))
