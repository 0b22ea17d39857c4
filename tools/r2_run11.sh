#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_gpu_msm.py -q -x > gpurun_out/r2_tests11.log 2>&1; echo "msm tests rc=$?"; tail -2 gpurun_out/r2_tests11.log
for P in 0 4 16; do
OZL_MSM_SCATTER_PARTS=$P timeout 300 python bench.py --steps 4 --warmup 2 --no-groth16 --no-ntt --strong-log-n 0 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('2^26 parts=$P', 'step', round(d['ms_per_step'],2), 'ms; e2e', round(d['e2e']['ms_per_step'],2), 'ms', d['verified_vs_known_dlog'], {k: round(v,2) for k,v in d['stages_ms'].items()})"
done
timeout 300 python bench.py --workload groth16 --no-cpu-baseline --concurrency 2 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('g16', round(d['ms_per_step'],3), 'ms', d['verified'], d['concurrent'], {k: round(v,2) for k,v in d['stages_ms'].items()})"
