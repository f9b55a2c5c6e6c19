"""Occupancy-weighted kernel time per frame pair from an ncu launch list (gpu__time_duration.sum per launch, --csv).

A launch's duration counts in full when its grid fills the machine and in proportion (CTAs x threads / resident threads) when it
does not: with 16 contexts pipelined the small-grid kernels (one-CTA bookkeeping, coarse pyramid levels) run beside other contexts'
kernels, the machine-filling ones do not.  The sum of the weighted times is the model of the pipelined step (bench.py `value`).

    python scripts/weighted_share.py profiles/r02b_launches_ncu.csv 5        # 5 = frame pairs in the capture
"""
import collections, csv, re, sys

path, pairs = sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
rows = list(csv.reader(open(path)))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
H = rows[hdr]
ki, vi, gi, bi = H.index("Kernel Name"), H.index("Metric Value"), H.index("Grid Size"), H.index("Block Size")


def prod(s):
    v = 1
    for x in re.findall(r"\d+", s):
        v *= int(x)
    return v


agg, tot, totw = collections.OrderedDict(), 0.0, 0.0
for r in rows[hdr + 1:]:
    if len(r) <= vi:
        continue
    name = r[ki].split("(")[0].replace("void ", "").replace("<unnamed>::", "").replace("akz::", "")
    t = float(r[vi].replace(",", "")) / 1000.0                       # us
    f = min(1.0, prod(r[gi]) * prod(r[bi]) / (148 * 2048 * 0.75))     # 0.75: typical residency of the register-limited kernels
    a = agg.setdefault(name, [0, 0.0, 0.0])
    a[0] += 1; a[1] += t; a[2] += t * f
    tot += t; totw += t * f
print(f"{'kernel':36s} {'launches':>9s} {'alone us':>10s} {'weighted us':>12s} {'share':>7s}   (per frame pair)")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][2]):
    if a[2] / pairs < 0.25:
        continue
    print(f"{k[:36]:36s} {a[0] / pairs:9.1f} {a[1] / pairs:10.1f} {a[2] / pairs:12.1f} {100 * a[2] / totw:6.1f}%")
print(f"{'total':36s} {'':9s} {tot / pairs:10.1f} {totw / pairs:12.1f}")
