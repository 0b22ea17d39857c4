#!/bin/bash
# One GPU box: parity suite, smoke, the three default bench lines.  Run under gpurun; results in gpurun_out/.
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3)
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/bench_msm.json 2> gpurun_out/bench_msm.err; tail -2 gpurun_out/bench_msm.err
timeout 600 python bench.py --workload groth16 > gpurun_out/bench_g16.json 2> gpurun_out/bench_g16.err
timeout 300 python bench.py --workload ntt > gpurun_out/bench_ntt.json 2> gpurun_out/bench_ntt.err
python - <<'EOF'
import json
d = json.load(open("gpurun_out/bench_msm.json"))
print("msm", d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["pipelined"]["value"], d["stages_ms"], d["fma_pipe"]["frac"],
      d["roofline"]["frac"], d["verified_vs_known_dlog"], d["gpu_launches"])
d = json.load(open("gpurun_out/bench_g16.json"))
print("g16", d["value"], d["ms_per_step"], d["concurrent"])
d = json.load(open("gpurun_out/bench_ntt.json"))
print("ntt", d["value"], d["ms_per_step"], d["config"]["workload"])
EOF
