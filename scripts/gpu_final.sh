#!/bin/bash
# Round-end check on one B200 (run under gpurun): GPU health, the whole `-m gpu` suite, then the default bench line.
# Usage: bash scripts/gpu_final.sh <tag>
cd ${GRAFT_REPO_ROOT:-.}
T=${1:-final}
nvidia-smi -L || { echo GPU_DEAD_AT_START; exit 0; }
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
nvidia-smi -L > /dev/null || { echo GPU_DEAD_AFTER_TESTS; exit 0; }
timeout 300 python bench.py > gpurun_out/${T}_bench_n1.json 2> gpurun_out/${T}.err
python - <<P
import json
d = json.loads(open("gpurun_out/${T}_bench_n1.json").read().strip().splitlines()[-1])
k = d["roofline"]["kernels"]
print(round(d["value"], 1), round(d["e2e"]["value"], 1), d["inliers_equal_oracle"], d["ransac_two_view"]["single_call_latency_ms"],
      {n: round(x["ms_per_pair"], 3) for n, x in k.items() if "score" in n}, d["cpu_baseline"]["value"], d["clocks"])
P
nvidia-smi -L > /dev/null && echo GPU_OK_AT_END
