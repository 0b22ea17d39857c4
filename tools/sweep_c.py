#!/usr/bin/env python3
"""Window-width / precompute-factor sweep (device-resident scalars, CUDA-event timing)."""
import sys, json, time
import numpy as np
import torch
sys.path.insert(0, ".")
import openzl_b200 as ozl
from bench import device_scalars, R381
R254 = 21888242871839275222246405745257275088548364400416034343698204186575808495617

def main():
    curve_name, log_n = sys.argv[1], int(sys.argv[2])
    cs = [int(x) for x in sys.argv[3].split(",")]
    fs = [int(x) for x in sys.argv[4].split(",")]
    curve = ozl.CURVE_IDS[curve_name]
    mod = R381 if "381" in curve_name else R254
    n = 1 << log_n
    dev = torch.device("cuda", 0)
    ctx = ozl.Context(0)
    limbs = ctx._lib.ozl_curve_coord_limbs(curve)
    sc = device_scalars(n, mod, 7, dev)
    out = torch.zeros(3 * limbs, dtype=torch.int64, device=dev)
    for c in cs:
        for f in fs:
            ctx.set_window_bits(c)
            b = ctx.generate_bases(curve, 1, n)
            if f > 1:
                b.precompute(f)
            ctx.enable_timing(True)
            ts = []
            for rep in range(4):
                b.msm_device(sc.data_ptr(), n, out.data_ptr())
                st = ctx.stage_times()
                ts.append(sum(ms for _, ms, _ in st))
            ctx.enable_timing(False)
            best = min(ts[1:])
            print(json.dumps({"curve": curve_name, "log_n": log_n, "c": c, "factor": f, "ms": round(best, 3),
                              "stages": {k: round(v, 2) for k, v, _ in st}}), flush=True)
            b.free()

if __name__ == "__main__":
    main()
