"""-m gpu: the PIPELINED MSM (runtime.cuh: bucket intervals accumulated on two side streams while the next interval is
still being sorted) returns the known-discrete-log answer.  Production enables it from ~2^24 points; here it is forced
at small sizes (OZL_MSM_PIPE = interval count, OZL_MSM_SCATTER_PARTS = bucket ranges per set), with and without shifted
base copies, for device-resident scalars and for host scalars that arrive in point-range batches, and with skewed
scalars that leave most intervals empty.  The switches are read once per process, so each case is its own interpreter.
The full-size run (2^24, production thresholds) is tests/test_gpu_msm.py::test_msm_known_dlog_large."""
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
name, log_n, factor, skew = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
R = {"bls": 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001,
     "bn2": 21888242871839275222246405745257275088548364400416034343698204186575808495617}[name[:3]]
n = 1 << log_n
ctx = ozl.Context(0)
h = ctx.generate_bases(ozl.CURVE_IDS[name], 3, n)
if factor > 1:
    h = h.precompute(factor)
s = random_scalars(n, R, seed=23)
if skew:
    s[: n // 2, 1:] = 0          # half the scalars below 2^64: their upper windows are empty
    s[n // 2 : n // 2 + n // 8] = s[0]
import torch
out = []
for _ in range(2):               # host scalars (ozl_msm; batches when OZL_MSM_H2D_SPLIT is set); the second call reuses
    aff, inf = ctx.jacobian_to_affine(ozl.CURVE_IDS[name], h.msm(s))   # the workspace, its side streams and events
    out.append([int(v) for v in aff] + [int(inf)])
d_s = torch.from_numpy(s.view(np.int64)).cuda()
limbs = 6 if name == "bls12_381_g1" else (4 if name == "bn254_g1" else (12 if name == "bls12_381_g2" else 8))
d_out = torch.zeros(3 * limbs, dtype=torch.int64, device="cuda")
torch.cuda.synchronize()
h.msm_device(d_s.data_ptr(), n, d_out.data_ptr())     # device-resident scalars: one batch
ctx.synchronize()
aff, inf = ctx.jacobian_to_affine(ozl.CURVE_IDS[name], d_out.cpu().numpy().view(np.uint64))
out.append([int(v) for v in aff] + [int(inf)])
print(json.dumps(out))
""" % ROOT


def _expected(name, log_n, skew):
    from oracle import cbind
    from tests.util import random_scalars
    r = {"bls": 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001,
         "bn2": 21888242871839275222246405745257275088548364400416034343698204186575808495617}[name[:3]]
    n = 1 << log_n
    s = random_scalars(n, r, seed=23)
    if skew:
        s[: n // 2, 1:] = 0
        s[n // 2: n // 2 + n // 8] = s[0]
    field = "bls12_381_fr" if name.startswith("bls") else "bn254_fr"
    k = cbind.dot_mod_r(field, s, np.arange(3, 3 + n, dtype=np.uint64))
    exp, _ = cbind.to_affine(name, cbind.gen_mul(name, k))
    return [int(v) for v in exp] + [0]


@pytest.mark.parametrize("name,log_n,factor,skew,env", [
    ("bls12_381_g1", 16, 1, 0, {"OZL_MSM_PIPE": "4", "OZL_MSM_SCATTER_PARTS": "4"}),
    ("bls12_381_g1", 16, 4, 0, {"OZL_MSM_PIPE": "8", "OZL_MSM_SCATTER_PARTS": "4"}),
    ("bls12_381_g1", 16, 32, 1, {"OZL_MSM_PIPE": "8", "OZL_MSM_SCATTER_PARTS": "8"}),
    ("bn254_g1", 15, 1, 1, {"OZL_MSM_PIPE": "16", "OZL_MSM_SCATTER_PARTS": "2"}),
    ("bn254_g1", 17, 32, 0, {"OZL_MSM_PIPE": "3", "OZL_MSM_SCATTER_PARTS": "16", "OZL_MSM_H2D_SPLIT": "0.1,0.3"}),
    ("bls12_381_g1", 17, 2, 0, {"OZL_MSM_PIPE": "5", "OZL_MSM_SCATTER_PARTS": "3", "OZL_MSM_H2D_SPLIT": "0.05,0.2,0.3"}),
    ("bls12_381_g2", 13, 1, 0, {"OZL_MSM_PIPE": "4", "OZL_MSM_SCATTER_PARTS": "4"}),   # G2 keeps the single launch
])
def test_pipelined_msm_matches_known_dlog(name, log_n, factor, skew, env):
    exp = _expected(name, log_n, skew)
    res = subprocess.run([sys.executable, "-c", SCRIPT, name, str(log_n), str(factor), str(skew)], capture_output=True, text=True,
                         env=dict(os.environ, **env), timeout=600)
    assert res.returncode == 0, res.stderr[-600:]
    got = json.loads(res.stdout.strip().splitlines()[-1])
    assert len(got) == 3
    for g in got:
        assert g == exp, (name, env)
