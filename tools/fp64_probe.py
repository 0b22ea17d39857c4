#!/usr/bin/env python3
"""Throughput of the field multiplier variants through ozl_bench_field_mul: integer (0), FP64 pipe (2), two integer +
two FP64 chains per thread (3).  Prints one JSON line."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import openzl_b200 as ozl
ctx = ozl.Context(0)
out = {}
for name, fid in (("int_381", 0), ("fp64_381", 2), ("mixed_2int_2fp64", 3), ("interleaved_pair", 4), ("int_254", 1)):
    ctx.bench_field_mul(fid, 500)
    out[name] = ctx.bench_field_mul(fid, 3000)
out["mixed_over_int"] = out["mixed_2int_2fp64"] / out["int_381"]
out["interleaved_over_int"] = out["interleaved_pair"] / out["int_381"]
print(json.dumps(out))
