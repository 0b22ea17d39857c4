#!/bin/bash
# GPU session 3: tile NTT parity + timing, ncu of the tile kernel, e2e batch-split sweep, Groth16 window sweep
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_gpu_ntt.py tests/test_gpu_groth16.py -q > gpurun_out/r2_ntt_tests.log 2>&1; echo "ntt tests rc=$?"; tail -3 gpurun_out/r2_ntt_tests.log
for L in 20 22 24; do
  for T in 1 0; do
    OZL_NTT_TILE=$T timeout 300 python bench.py --workload ntt --log-n $L --steps 20 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('ntt', $L, 'tile=$T', round(d['ms_per_step'],4), 'ms', d['config']['workload'][-16:], 'frac', round(d['fma_pipe']['frac'],3), d['verified_round_trip'])"
  done
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_ntt_tile -c 3 -o gpurun_out/r2_ncu_ntt_tile -f python bench.py --workload ntt --log-n 24 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r2_ncu_ntt.log 2>&1; echo "ncu rc=$?"
for S in "0.125" "0.0625,0.4375" "0.03125,0.21875" "0.03125,0.125,0.35"; do
  OZL_MSM_H2D_SPLIT=$S timeout 300 python bench.py --steps 4 --warmup 2 --no-groth16 --no-ntt --strong-log-n 0 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('split $S', 'value', round(d['ms_per_step'],2), 'ms; e2e', round(d['e2e']['ms_per_step'],2), 'ms', d['verified_vs_known_dlog'])"
done
for C in 0 17 18 20; do
  timeout 300 python bench.py --workload groth16 --window-bits $C --no-cpu-baseline --concurrency 1 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('g16 c=$C', round(d['ms_per_step'],3), 'ms', d['verified'], {k: round(v,2) for k,v in d['stages_ms'].items()})"
done
