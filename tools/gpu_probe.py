#!/usr/bin/env python3
"""First-contact GPU probe: field-mul micro-benchmark, MSM stage timings, NTT timings."""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import openzl_b200 as ozl
from tests.util import random_scalars

R381 = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
R254 = 21888242871839275222246405745257275088548364400416034343698204186575808495617


def main():
    ctx = ozl.Context(0)
    out = {}
    for fid, nm in ((0, "fq381"), (1, "fq254")):
        out[f"mul_per_s_{nm}"] = ctx.bench_field_mul(fid, 4000)
    print(json.dumps(out), flush=True)
    logs = [int(a) for a in sys.argv[1:]] or [16, 20, 22]
    for log_n in logs:
        n = 1 << log_n
        t0 = time.time()
        h = ctx.generate_bases(ozl.BLS12_381_G1, 1, n)
        t1 = time.time()
        s = random_scalars(n, R381, seed=log_n)
        ctx.enable_timing(True)
        for rep in range(2):
            t2 = time.time()
            h.msm(s)
            t3 = time.time()
            st = ctx.stage_times()
        print(json.dumps({"log_n": log_n, "c": ctx.window_bits(0, n), "gen_s": t1 - t0, "msm_wall_s": t3 - t2,
                          "pts_per_s": n / (t3 - t2), "stages": st}), flush=True)
        ctx.enable_timing(False)
        h.free()
    for log_n in (20, 24):
        x = random_scalars(1 << log_n, R254, seed=1)
        ctx.enable_timing(True)
        for rep in range(2):
            t0 = time.time()
            ctx.ntt(ozl.BN254_FR, x)
            t1 = time.time()
            st = ctx.stage_times()
        print(json.dumps({"ntt_log_n": log_n, "wall_s": t1 - t0, "stages": st}), flush=True)
        ctx.enable_timing(False)


if __name__ == "__main__":
    main()
