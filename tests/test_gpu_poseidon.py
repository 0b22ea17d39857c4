"""-m gpu: the reference's own known-answer vector driven through the CUDA Montgomery multiplier.

SURVEY.md section 8c: the only golden vectors the reference holds for this path pin BLS12-381 Fr
arithmetic (Poseidon width-3 KAT at /root/reference/openzl-tutorials/src/poseidon.rs:388-401, the
189 Grain-LFSR round constants and the Cauchy MDS matrices in plugins/arkworks/src/poseidon/).
`ozl_fr_poseidon_permute` runs the permutation of /root/reference/openzl-crypto/src/poseidon/mod.rs:156-283
on the device -- 63 rounds of field additions, x^5 S-boxes and 3x3 MDS products in Montgomery form --
so the fixture in tests/golden/ (extracted from the reference by tests/golden/make_golden.py) is the
one place a reference-held vector touches a CUDA kernel directly, not through the oracle."""
import json
import os

import numpy as np
import pytest

import openzl_b200 as ozl
from openzl_b200.groth16 import ints_to_limbs, limbs_to_ints
from oracle import fields, poseidon as oposeidon

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _golden():
    with open(os.path.join(HERE, "golden", "poseidon_bls12_381_fr.json")) as fh:
        return json.load(fh)


def _permute(ctx, field_id, p, states, width, rf, rp, keys, mds):
    st = ints_to_limbs([v for s in states for v in s], p, mont=True)
    k = ints_to_limbs(keys, p, mont=True)
    m = ints_to_limbs([v for row in mds for v in row], p, mont=True)
    ctx._check(ctx._lib.ozl_fr_poseidon_permute(ctx._h, field_id, st.ctypes.data, len(states), width, rf, rp,
                                                k.ctypes.data, m.ctypes.data), "ozl_fr_poseidon_permute")
    flat = limbs_to_ints(st, p, mont=True)
    return [flat[i * width:(i + 1) * width] for i in range(len(states))]


def test_device_poseidon_width3_kat_from_reference_goldens(ctx):
    g = _golden()
    p = fields.BLS12_381_FR.p
    kat = g["permutation_width3"]
    lf = g["lfsr_values"]
    assert (lf["width"], lf["full_rounds"], lf["partial_rounds"]) == (3, 8, 55)
    keys = [int(x) for x in lf["values"]]
    flat = [int(x) for x in g["mds"]["3"]]
    mds = [flat[0:3], flat[3:6], flat[6:9]]
    inp = [int(x) for x in kat["input"]]
    exp = [int(x) for x in kat["tutorial_output"]]
    assert len(keys) == 189 and inp == [3, 1, 2]
    # every input element, round key and MDS entry comes from the reference's files; nothing from the oracle
    out = _permute(ctx, ozl.BLS12_381_FR, p, [inp] * 70 + [[0, 0, 0]], 3, 8, 55, keys, mds)
    for row in out[:70]:                                  # more states than one warp: every lane agrees
        assert row == exp
    assert out[70] != exp


@pytest.mark.parametrize("fname,fid", [("bn254_fr", 0), ("bls12_381_fr", 1)])
@pytest.mark.parametrize("width", [2, 3, 5, 12])
def test_device_poseidon_matches_oracle(ctx, fname, fid, width):
    """Other widths / the BN254 field (the Groth16 workload's): device permutation == big-int oracle."""
    import random
    f = fields.FIELDS[fname]
    rf, rp = 8, 11
    keys = oposeidon.generate_round_constants(f, width, rf, rp)
    mds = oposeidon.generate_mds(f, width)
    rnd = random.Random(width)
    states = [[rnd.randrange(f.p) for _ in range(width)] for _ in range(33)] + [[0] * width, [f.p - 1] * width]
    got = _permute(ctx, fid, f.p, states, width, rf, rp, keys, mds)
    for s, gts in zip(states, got):
        assert gts == oposeidon.permute(f, s, keys, mds, rf, rp)
