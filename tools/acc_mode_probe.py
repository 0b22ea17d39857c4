#!/usr/bin/env python3
"""Times one MSM configuration (curve, log n, precompute) with device-resident scalars; the accumulate
variant comes from OZL_ACC_MODE in the environment (read once per process).  Prints one JSON line."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import openzl_b200 as ozl
from bench import device_scalars

R = {"bls12_381": 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001,
     "bn254": 21888242871839275222246405745257275088548364400416034343698204186575808495617}
name, log_n = sys.argv[1], int(sys.argv[2])
pre = int(sys.argv[3]) if len(sys.argv) > 3 else 32
n = 1 << log_n
ctx = ozl.Context(0)
ctx.use_torch_stream()
curve = ozl.CURVE_IDS[name]
h = ctx.generate_bases(curve, 1, n)
if pre > 1:
    h.precompute(pre)
dev = torch.device("cuda", 0)
sc = device_scalars(n, R["bls12_381" if name.startswith("bls") else "bn254"], 7, dev)
limbs = ctx._lib.ozl_curve_coord_limbs(curve)
out = torch.zeros(3 * limbs, dtype=torch.int64, device=dev)
for _ in range(2):
    h.msm_device(sc.data_ptr(), n, out.data_ptr())
torch.cuda.synchronize()
ctx.enable_timing(True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
reps = 4
acc = []
for _ in range(reps):
    h.msm_device(sc.data_ptr(), n, out.data_ptr())
    acc.append(dict((k, v) for k, v, _ in ctx.stage_times()).get("accumulate"))
e1.record()
torch.cuda.synchronize()
aff, _ = ctx.jacobian_to_affine(curve, out.cpu().numpy().view(np.uint64))
print(json.dumps({"curve": name, "log_n": log_n, "mode": os.environ.get("OZL_ACC_MODE", "default"), "ms": e0.elapsed_time(e1) / reps,
                  "accumulate_ms": float(np.mean(acc)), "info": h.info(n), "x0": int(aff[0])}))
