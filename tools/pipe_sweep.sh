#!/bin/bash
# A/B runs of the pipelined MSM at 2^26: "name ENV=..." per line
run() {
  name=$1; shift
  env "$@" python bench.py --steps 3 --warmup 2 --no-groth16 --no-ntt --strong-log-n 0 --no-cpu-baseline > gpurun_out/ps_$name.json 2> gpurun_out/ps_$name.err
  python - "$name" <<PY
import json,sys
n=sys.argv[1]
try:
    d=json.load(open(f"gpurun_out/ps_{n}.json"))
    st={k:round(v,2) for k,v in d["stages_ms"].items()}
    print(n, round(d["ms_per_step"],2), "e2e", round(d["e2e"]["ms_per_step"],2), st, d["verified_vs_known_dlog"])
except Exception as e: print(n, "ERR", e)
PY
}
