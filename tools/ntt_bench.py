#!/usr/bin/env python3
"""NTT sweep (BASELINE config 3): BN254 Fr, log n = 20..24, device-resident, CUDA-event timing."""
import json, sys
import numpy as np
import torch
sys.path.insert(0, ".")
import openzl_b200 as ozl
from bench import device_scalars
R254 = 21888242871839275222246405745257275088548364400416034343698204186575808495617
R381 = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001

def main():
    logs = [int(x) for x in (sys.argv[1].split(",") if len(sys.argv) > 1 else "20,22,24".split(","))]
    dev = torch.device("cuda", 0)
    ctx = ozl.Context(0)
    ctx.use_torch_stream()
    for field, fid, mod in (("bn254_fr", ozl.BN254_FR, R254), ("bls12_381_fr", ozl.BLS12_381_FR, R381)):
        for log_n in logs:
            n = 1 << log_n
            x = device_scalars(n, mod, 3, dev)
            for inverse, coset in ((False, False), (True, True)):
                for _ in range(3):
                    ctx.ntt_device(fid, x.data_ptr(), log_n, inverse, coset)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                reps = 10
                e0.record()
                for _ in range(reps):
                    ctx.ntt_device(fid, x.data_ptr(), log_n, inverse, coset)
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / reps
                passes = (log_n + 2) // 3
                print(json.dumps({"field": field, "log_n": log_n, "inverse": inverse, "coset": coset, "ms": round(ms, 4),
                                  "elements_per_s": n / (ms * 1e-3), "algorithmic_GBps": n * 64 / (ms * 1e-3) / 1e9,
                                  "butterfly_mul_per_s": (n / 2) * log_n / (ms * 1e-3), "passes": passes}), flush=True)
        if len(sys.argv) > 2 and sys.argv[2] == "one":
            break

if __name__ == "__main__":
    main()
