#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
: > gpurun_out/r2_acc_modes.jsonl
for cfg in "bls12_381_g1 24" "bls12_381_g2 21" "bn254_g1 22" "bn254_g2 21"; do
  for M in 0 1 2 3; do
    OZL_ACC_MODE=$M timeout 300 python tools/acc_mode_probe.py $cfg 2>/dev/null | tee -a gpurun_out/r2_acc_modes.jsonl | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['curve'], d['log_n'], 'mode', d['mode'], round(d['ms'],3), 'ms acc', round(d['accumulate_ms'],3), hex(d['x0'])[:10])"
  done
done
OZL_ACC_MODE=3 timeout 300 python bench.py --steps 4 --warmup 2 --no-groth16 --no-ntt --strong-log-n 0 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('2^26 acc mode 3', 'step', round(d['ms_per_step'],2), 'ms; e2e', round(d['e2e']['ms_per_step'],2), 'ms', d['verified_vs_known_dlog'], {k: round(v,2) for k,v in d['stages_ms'].items()})"
for L in 20 24; do
  timeout 300 python bench.py --workload ntt --log-n $L --steps 20 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('ntt', $L, round(d['ms_per_step'],4), 'ms', 'frac', round(d['fma_pipe']['frac'],3), d['verified_round_trip'])"
done
