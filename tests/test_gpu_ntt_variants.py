"""-m gpu: every tile variant of the NTT (OZL_NTT_R4 = 2: radix-4 rounds, four CTAs per SM, the shipped one; 1: three
CTAs; 0: the radix-8 rounds on 2048-element tiles; OZL_NTT_PAIRED=1: radix-8 with the paired out-of-line multiplier;
OZL_NTT_TILE=0: register-only passes) returns the oracle's transform bit for bit, in all four modes, on both fields, at
sizes that exercise every pass plan (odd and even log n, one to three tile passes).  The switches are read once per
process, so each variant runs in its own interpreter."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r"""
import json, sys
sys.path.insert(0, %r)
import numpy as np
import openzl_b200 as ozl
from oracle import cbind, fields
from tests.util import random_scalars
ctx = ozl.Context(0)
bad = []
for fname, fid in (("bn254_fr", ozl.BN254_FR), ("bls12_381_fr", ozl.BLS12_381_FR)):
    p = fields.FIELDS[fname].p
    for log_n in (9, 10, 11, 12, 13, 15, 16, 17, 18):
        x = random_scalars(1 << log_n, p, seed=100 + log_n)
        for inverse, coset in ((False, False), (True, False), (False, True), (True, True)):
            got = x.copy()
            ctx.ntt(fid, got, inverse=inverse, coset=coset)
            if not (got == cbind.ntt(fname, x, inverse=inverse, coset=coset)).all():
                bad.append([fname, log_n, inverse, coset])
print(json.dumps(bad))
""" % ROOT


@pytest.mark.parametrize("env", [
    {"OZL_NTT_R4": "2"},
    {"OZL_NTT_R4": "1"},
    {"OZL_NTT_R4": "0"},
    {"OZL_NTT_R4": "0", "OZL_NTT_PAIRED": "1"},
    {"OZL_NTT_TILE": "0"},
])
def test_ntt_tile_variants_match_oracle(env):
    res = subprocess.run([sys.executable, "-c", SCRIPT], capture_output=True, text=True, env=dict(os.environ, **env), timeout=600)
    assert res.returncode == 0, res.stderr[-600:]
    assert json.loads(res.stdout.strip().splitlines()[-1]) == [], env
