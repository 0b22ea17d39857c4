#!/usr/bin/env python3
"""For each size, sweep the window width with a FULL set of shifted base copies (factor = number of
windows, one bucket set) and with the planner's default (4 copies); device-resident scalars."""
import sys, json
import torch
sys.path.insert(0, ".")
import openzl_b200 as ozl
from bench import device_scalars, R381


def main():
    curve = ozl.BLS12_381_G1
    dev = torch.device("cuda", 0)
    ctx = ozl.Context(0)
    out = torch.zeros(18, dtype=torch.int64, device=dev)
    plan = {20: [15, 16, 17, 18, 19, 20], 22: [16, 17, 18, 19, 20, 21, 22], 24: [18, 19, 20, 21, 22, 23], 26: [20, 21, 22, 23, 24]}
    only = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else sorted(plan)
    for log_n in only:
        n = 1 << log_n
        sc = device_scalars(n, R381, 7, dev)
        configs = [(0, 4)] + [(c, (256 + c - 1) // c) for c in plan[log_n]]
        for c, f in configs:
            if n * 96 * f > 90e9:
                continue
            ctx.set_window_bits(c)
            b = ctx.generate_bases(curve, 1, n)
            b.precompute(f)
            ctx.enable_timing(True)
            ts = []
            for rep in range(4):
                b.msm_device(sc.data_ptr(), n, out.data_ptr())
                st = ctx.stage_times()
                ts.append(sum(ms for _, ms, _ in st))
            ctx.enable_timing(False)
            info = b.info(n)
            print(json.dumps({"log_n": log_n, "c": info["c"], "factor": f, "bucket_sets": info["bucket_sets"], "ms": round(min(ts[1:]), 3),
                              "pts_per_s": round(n / (min(ts[1:]) * 1e-3)), "stages": {k: round(v, 2) for k, v, _ in st}}), flush=True)
            b.free()
    ctx.set_window_bits(0)


if __name__ == "__main__":
    main()
