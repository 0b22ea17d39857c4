#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2_gpu_tests.log 2>&1; echo "gpu tests rc=$?"; tail -3 gpurun_out/r2_gpu_tests.log
for cfg in "bn254_g1 22" "bls12_381_g1 24"; do
 for M in 0 5 6; do
  OZL_ACC_MODE=$M timeout 300 python tools/acc_mode_probe.py $cfg 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['curve'], d['log_n'], 'mode', d['mode'], round(d['ms'],3), 'ms acc', round(d['accumulate_ms'],3), hex(d['x0'])[:10])"
 done
done
for M in 0 5 6; do
  OZL_ACC_MODE=$M timeout 300 python bench.py --workload groth16 --no-cpu-baseline --concurrency 2 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('g16 acc mode $M', round(d['ms_per_step'],3), 'ms', d['verified'], d['concurrent'], {k: round(v,2) for k,v in d['stages_ms'].items()})"
done
