"""Pin the oracle's field arithmetic against the reference's OWN golden vectors
(tests/golden/poseidon_bls12_381_fr.json, extracted by tests/golden/make_golden.py from
/root/reference/plugins/arkworks/src/poseidon/{parameters_hardcoded_test,mds_hardcoded_tests,
permutation_hardcoded_test} and /root/reference/openzl-tutorials/src/poseidon.rs:388-401).

These are the only fixed vectors the reference holds for arithmetic on this path; the reference
has no MSM / NTT / proof vectors, so MSM/NTT parity is "unpinned" there
and pinned by the mathematical cross-checks in test_oracle_math.py.
"""
import json
import os

import pytest

from oracle import curves, fields, poseidon

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "poseidon_bls12_381_fr.json")))
F = fields.BLS12_381_FR


def test_constants_self_check():
    fields.self_check()
    curves.self_check()


def test_lfsr_round_constants_match_reference():
    g = GOLD["lfsr_values"]
    got = poseidon.generate_round_constants(F, g["width"], g["full_rounds"], g["partial_rounds"])
    assert [str(v) for v in got] == g["values"]
    assert len(got) == 189


@pytest.mark.parametrize("t", range(2, 13))
def test_mds_matches_reference(t):
    m = poseidon.generate_mds(F, t)
    flat = [str(v) for row in m for v in row]
    assert flat == GOLD["mds"][str(t)]


def test_permutation_kat_matches_reference():
    g = GOLD["permutation_width3"]
    rk = poseidon.generate_round_constants(F, 3, 8, 55)
    mds = poseidon.generate_mds(F, 3)
    out = poseidon.permute(F, [int(v) for v in g["input"]], rk, mds, 8, 55)
    assert [str(v) for v in out] == g["output"]
    assert g["output"] == g["tutorial_output"]


def test_kat_through_montgomery_form():
    """Same KAT with every product done as a Montgomery product, the form the kernels use."""
    g = GOLD["permutation_width3"]
    rk = poseidon.generate_round_constants(F, 3, 8, 55)
    mds = poseidon.generate_mds(F, 3)
    M = F.to_mont

    def mmul(a, b):
        return F.mont_mul(a, b)

    state = [M(int(v)) for v in g["input"]]
    rkm = [M(v) for v in rk]
    mdsm = [[M(v) for v in row] for row in mds]
    for rnd in range(63):
        state = [(s + k) % F.p for s, k in zip(state, rkm[rnd * 3:(rnd + 1) * 3])]
        idx = range(3) if (rnd < 4 or rnd >= 59) else [0]
        for i in idx:
            s2 = mmul(state[i], state[i])
            state[i] = mmul(mmul(s2, s2), state[i])
        state = [sum(mmul(mdsm[i][j], state[j]) for j in range(3)) % F.p for i in range(3)]
    assert [str(F.from_mont(s)) for s in state] == g["output"]
