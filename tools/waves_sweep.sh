#!/bin/bash
# A/B of the wave-aware slice length (OZL_MSM_WAVES=1 default / 0 = round-1 rule) over the MSM sweep sizes and one Groth16 proof
for w in 1 0; do
  for ln in 20 22 24; do
    OZL_MSM_WAVES=$w python bench.py --log-n $ln --steps 5 --warmup 3 --no-groth16 --no-ntt --strong-log-n 0 --no-cpu-baseline > gpurun_out/wv_${w}_$ln.json 2> gpurun_out/wv_${w}_$ln.err
  done
  OZL_MSM_WAVES=$w python bench.py --workload groth16 --no-cpu-baseline > gpurun_out/wv_${w}_g16.json 2> gpurun_out/wv_${w}_g16.err
done
python - <<PY
import json
for w in (1,0):
    for ln in (20,22,24):
        try:
            d=json.load(open(f"gpurun_out/wv_{w}_{ln}.json"))
            print(w, ln, round(d["ms_per_step"],3), "e2e", round(d["e2e"]["ms_per_step"],3), {k:round(v,2) for k,v in d["stages_ms"].items()}, d["verified_vs_known_dlog"])
        except Exception as e: print(w, ln, "ERR", e)
    try:
        d=json.load(open(f"gpurun_out/wv_{w}_g16.json"))
        print(w, "g16", round(d["ms_per_step"],3), d.get("concurrent"), {k:round(v,2) for k,v in d["stages_ms"].items()}, d["verified"])
    except Exception as e: print(w, "g16 ERR", e)
PY
