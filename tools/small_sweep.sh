#!/bin/bash
# A/B of accumulation grid (persistent / yielding) and slice length at the sweep sizes: "run NAME LOGN ENV=..."
run() {
  name=$1; ln=$2; shift; shift
  env "$@" python bench.py --log-n $ln --steps 5 --warmup 3 --no-groth16 --no-ntt --strong-log-n 0 --no-cpu-baseline > gpurun_out/sm_${name}_$ln.json 2> gpurun_out/sm_${name}_$ln.err
  python - "$name" "$ln" <<PY
import json,sys
n,ln=sys.argv[1],sys.argv[2]
try:
    d=json.load(open(f"gpurun_out/sm_{n}_{ln}.json"))
    print(n, ln, round(d["ms_per_step"],3), {k:round(v,2) for k,v in d["stages_ms"].items()}, d["verified_vs_known_dlog"])
except Exception as e: print(n, ln, "ERR", e)
PY
}
