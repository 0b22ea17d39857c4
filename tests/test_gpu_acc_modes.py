"""-m gpu: every variant of the hot loop's field products (OZL_ACC_MODE 0..7: inlined, out-of-line multiplier,
paired, dedicated squaring, Karatsuba + separated reduction, fused dual product, inlined + fused, FP64-pipe
products; 8 / 18: lazily reduced Fq2 products on the G2 curves at either CTA count; 15: the other CTA count) returns the same point, equal to the known-discrete-log answer.  The mode is read once per process,
so each variant runs in its own interpreter.  The multiplier variants themselves are checked limb by limb on the
CPU (tests/test_host_emu.py); this is their device side."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r"""
import json, os, sys
sys.path.insert(0, %r)
import numpy as np
import openzl_b200 as ozl
from tests.util import random_scalars
name, log_n = sys.argv[1], int(sys.argv[2])
R = {"bls": 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001,
     "bn2": 21888242871839275222246405745257275088548364400416034343698204186575808495617}[name[:3]]
n = 1 << log_n
ctx = ozl.Context(0)
h = ctx.generate_bases(ozl.CURVE_IDS[name], 3, n).precompute(4)
s = random_scalars(n, R, seed=11)
aff, inf = ctx.jacobian_to_affine(ozl.CURVE_IDS[name], h.msm(s))
print(json.dumps({"aff": [int(v) for v in aff], "inf": bool(inf)}))
""" % ROOT


@pytest.mark.parametrize("name,log_n", [("bls12_381_g1", 16), ("bn254_g1", 15), ("bls12_381_g2", 13), ("bn254_g2", 13)])
def test_accumulate_variants_agree(name, log_n):
    from oracle import cbind
    from tests.util import random_scalars
    r = {"bls": 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001,
         "bn2": 21888242871839275222246405745257275088548364400416034343698204186575808495617}[name[:3]]
    n = 1 << log_n
    field = "bls12_381_fr" if name.startswith("bls") else "bn254_fr"
    k = cbind.dot_mod_r(field, random_scalars(n, r, seed=11), np.arange(3, 3 + n, dtype=np.uint64))
    exp, _ = cbind.to_affine(name, cbind.gen_mul(name, k))
    for mode in list(range(8)) + ([8, 15, 18] if name.endswith("g2") else []):
        env = dict(os.environ, OZL_ACC_MODE=str(mode))
        res = subprocess.run([sys.executable, "-c", SCRIPT, name, str(log_n)], capture_output=True, text=True, env=env, timeout=300)
        assert res.returncode == 0, (mode, res.stderr[-400:])
        got = json.loads(res.stdout.strip().splitlines()[-1])
        assert not got["inf"] and got["aff"] == [int(v) for v in exp], (name, mode)
