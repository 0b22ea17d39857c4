#!/bin/bash
# GPU session 5: accumulate variants (inline / out-of-line multiplier), collapse fix, NTT twiddle prefetch
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_gpu_msm.py tests/test_gpu_ntt.py -q -x > gpurun_out/r2_tests5.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/r2_tests5.log
for M in 0 1 2; do
  OZL_ACC_MODE=$M timeout 300 python bench.py --steps 4 --warmup 2 --no-groth16 --no-ntt --strong-log-n 0 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('acc mode $M', 'step', round(d['ms_per_step'],2), 'ms; e2e', round(d['e2e']['ms_per_step'],2), 'ms', d['verified_vs_known_dlog'], {k: round(v,2) for k,v in d['stages_ms'].items()})"
done
for L in 20 24; do
  timeout 300 python bench.py --workload ntt --log-n $L --steps 20 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('ntt', $L, round(d['ms_per_step'],4), 'ms', 'frac', round(d['fma_pipe']['frac'],3), d['verified_round_trip'])"
done
for M in 0 1; do
  OZL_ACC_MODE=$M timeout 300 python bench.py --workload groth16 --no-cpu-baseline --concurrency 1 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('g16 acc mode $M', round(d['ms_per_step'],3), 'ms', d['verified'], {k: round(v,2) for k,v in d['stages_ms'].items()})"
done
