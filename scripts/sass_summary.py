"""Counts of the Blackwell-native SASS mnemonics per kernel of cv_b200/libcvb200.so -> profiles/r02_sass_blackwell.txt"""
import collections, os, re, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = subprocess.run(["cuobjdump", "-sass", os.path.join(ROOT, "cv_b200", "libcvb200.so")], capture_output=True, text=True).stdout
cur, counts = None, collections.defaultdict(collections.Counter)
pat = re.compile(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)")
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1); continue
    m = pat.match(line)
    if m and cur:
        full = m.group(1).rstrip("."); op = full.split(".")[0]
        if op in ("UTCIMMA", "LDTM", "UTMALDG", "UBLKCP", "UTCBAR", "IMMA", "SYNCS"):
            counts[cur][full if op in ("UTMALDG", "LDTM", "UBLKCP", "IMMA") else op] += 1


def dem(n):
    d = subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip().replace("(anonymous namespace)::", "")
    d = re.sub(r"^void ", "", d)
    return d[:d.index("(")] if "(" in d else d


with open(os.path.join(ROOT, "profiles", "r02_sass_blackwell.txt"), "w") as f:
    f.write("# SASS mnemonics of the shipped cv_b200/libcvb200.so (cuobjdump -sass, sm_100a), per kernel: the Blackwell-native instructions\n"
            "# UTCIMMA = tcgen05.mma kind::i8, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit, UTMALDG = cp.async.bulk.tensor, UBLKCP = cp.async.bulk,\n"
            "# SYNCS = mbarrier ops; IMMA = legacy mma.sync (kept for A/B).  Regenerate: python scripts/sass_summary.py\n")
    for k, c in sorted(counts.items(), key=lambda kv: dem(kv[0])):
        f.write(f"{dem(k)}: " + ", ".join(f"{x} x{c[x]}" for x in sorted(c)) + "\n")
