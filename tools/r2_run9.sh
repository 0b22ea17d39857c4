#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
for cfg in "bls12_381_g1 24" "bn254_g1 22"; do
 for M in 0 3 5; do
  OZL_ACC_MODE=$M timeout 300 python tools/acc_mode_probe.py $cfg 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['curve'], d['log_n'], 'mode', d['mode'], round(d['ms'],3), 'ms acc', round(d['accumulate_ms'],3), hex(d['x0'])[:10])"
 done
done
OZL_ACC_MODE=5 timeout 300 python -m pytest tests/test_gpu_msm.py -q -x -k "parity or known_dlog or heavy or ragged or precomp" 2>&1 | tail -2
OZL_ACC_MODE=5 timeout 300 python bench.py --steps 4 --warmup 2 --no-groth16 --no-ntt --strong-log-n 0 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('2^26 acc mode 5', 'step', round(d['ms_per_step'],2), 'ms; e2e', round(d['e2e']['ms_per_step'],2), 'ms', d['verified_vs_known_dlog'], {k: round(v,2) for k,v in d['stages_ms'].items()})"
