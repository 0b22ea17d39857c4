#!/bin/bash
# GPU session 4: full -m gpu suite, padded tile NTT timing, ncu launch lists (MSM 2^26, Groth16), ncu --set full of k_accumulate_tma
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2_gpu_tests.log 2>&1; echo "gpu tests rc=$?"; tail -4 gpurun_out/r2_gpu_tests.log
for L in 20 22 24; do
  for T in 1 0; do
    OZL_NTT_TILE=$T timeout 300 python bench.py --workload ntt --log-n $L --steps 20 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('ntt', $L, 'tile=$T', round(d['ms_per_step'],4), 'ms', d['config']['workload'][-16:], 'frac', round(d['fma_pipe']['frac'],3), d['verified_round_trip'])"
  done
done
MSM="--steps 2 --warmup 1 --no-groth16 --no-ntt --strong-log-n 0 --no-cpu-baseline --no-verify"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r2_launches_msm_2p26.csv python bench.py $MSM > gpurun_out/r2_ncu_msm.log 2>&1; echo "ncu msm list rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2_launches_groth16.csv python bench.py --workload groth16 --g16-steps 1 --warmup 1 --concurrency 1 --no-cpu-baseline --no-verify > gpurun_out/r2_ncu_g16.log 2>&1; echo "ncu g16 list rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_accumulate_tma -s 4 -c 1 -o gpurun_out/r2_ncu_k_accumulate_2p26 -f python bench.py $MSM > gpurun_out/r2_ncu_acc.log 2>&1; echo "ncu acc full rc=$?"
