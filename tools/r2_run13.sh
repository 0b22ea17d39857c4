#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
for C in 14 15 16; do
  timeout 300 python bench.py --workload groth16 --window-bits $C --no-cpu-baseline --concurrency 2 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('g16 c=$C', round(d['ms_per_step'],3), 'ms', d['verified'], d['concurrent'], {k: round(v,2) for k,v in d['stages_ms'].items()})"
done
MSM="--steps 2 --warmup 1 --no-groth16 --no-ntt --strong-log-n 0 --no-cpu-baseline --no-verify"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_accumulate_tma -s 4 -c 1 -o gpurun_out/r2_ncu_k_accumulate_2p26_mode5 -f python bench.py $MSM > gpurun_out/r2_ncu_acc5.log 2>&1; echo "ncu acc full rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r2c_launches_msm_2p26.csv python bench.py $MSM > gpurun_out/r2_ncu_msm.log 2>&1; echo "ncu msm list rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2c_launches_groth16.csv python bench.py --workload groth16 --g16-steps 1 --warmup 1 --concurrency 1 --no-cpu-baseline --no-verify > gpurun_out/r2_ncu_g16.log 2>&1; echo "ncu g16 list rc=$?"
timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_msm.py -q -x -k "parity or ragged or heavy or zero_zero or non_canonical or precomputed" > gpurun_out/r2_sanitizer_msm.log 2>&1; echo "sanitizer msm rc=$?"; tail -3 gpurun_out/r2_sanitizer_msm.log
timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_ntt.py tests/test_gpu_poseidon.py -q -x -k "not 2p24 and not large" > gpurun_out/r2_sanitizer_ntt.log 2>&1; echo "sanitizer ntt rc=$?"; tail -3 gpurun_out/r2_sanitizer_ntt.log
