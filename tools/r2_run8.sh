#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
for M in 3 13; do
  OZL_ACC_MODE=$M timeout 300 python tools/acc_mode_probe.py bls12_381_g1 24 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['curve'], d['log_n'], 'mode', d['mode'], round(d['ms'],3), 'ms acc', round(d['accumulate_ms'],3), hex(d['x0'])[:10])"
done
MSM="--steps 2 --warmup 1 --no-groth16 --no-ntt --strong-log-n 0 --no-cpu-baseline --no-verify"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_accumulate_tma -s 4 -c 1 -o gpurun_out/r2_ncu_k_accumulate_2p26_mode3 -f python bench.py $MSM > gpurun_out/r2_ncu_acc3.log 2>&1; echo "ncu acc full rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r2b_launches_msm_2p26.csv python bench.py $MSM > gpurun_out/r2_ncu_msm.log 2>&1; echo "ncu msm list rc=$?"
