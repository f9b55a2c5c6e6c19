"""Pretty-print a bench.py JSON line (per-kernel breakdown)."""
import json
import sys

d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print({k: d.get(k) for k in ("value", "ms_per_step", "gpu_launches", "hamming_Gcmp_per_s")}, "e2e", d.get("e2e"))
print("cpu", d.get("cpu_baseline"), "clocks", d.get("clocks"))
r = d.get("roofline") or {}
print({k: r.get(k) for k in ("kernel", "achieved", "frac", "pipeline_alg_GBps", "pipeline_frac", "top_kernel_by_time")})
for k, v in sorted((r.get("kernels") or {}).items(), key=lambda kv: -kv[1]["ms_per_step"]):
    print(f"{k:20s} {v['launches_per_step']:5.0f} {v['ms_per_step'] * 1e3:9.1f} us {v['share'] * 100:5.1f}% {v['alg_GBps']}")
