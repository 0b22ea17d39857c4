#!/bin/bash
# A/B runs of the Groth16 prover's scheduling switches: "run NAME ENV=..." ; prints ms/proof, two-prover throughput, timeline
run() {
  EXTRA=${EXTRA:-}
  name=$1; shift
  env "$@" python bench.py --workload groth16 --no-cpu-baseline $EXTRA > gpurun_out/g16_$name.json 2> gpurun_out/g16_$name.err
  python - "$name" "$TL" <<PY
import json,sys
n=sys.argv[1]
try:
    d=json.load(open(f"gpurun_out/g16_{n}.json"))
    print(n, round(d["ms_per_step"],3), d["concurrent"], d["verified"])
    if sys.argv[2]=="1":
        for nm,a,b in d["timeline_ms"]: print(f"   {a:8.3f} {b:8.3f} {b-a:7.3f}  {nm}")
except Exception as e: print(n, "ERR", e)
PY
}
