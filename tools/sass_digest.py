"""SASS digest of libchd_b200.so: per kernel, how many tensor / TMA / async-copy / barrier instructions it contains
(cuobjdump -sass; no GPU needed).  python tools/sass_digest.py > profiles/r02_sass_digest.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "computationalhypergraphdiscovery_b200", "libchd_b200.so")
WANT = ["DMMA", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "LDGSTS", "LDG", "STG", "LDS", "STS", "BAR", "UCGABAR",
        "DFMA", "DADD", "DMUL", "SHFL", "ATOM", "RED", "MEMBAR"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    counts, cur = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = name.replace("(anonymous namespace)::", "").replace("chd::", "")
            cur = re.sub(r"^void ", "", re.sub(r"\(.*", "", name))
            counts[cur] = collections.Counter()
            continue
        m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m and cur:
            counts[cur][m.group(1)] += 1
    print("# SASS digest of libchd_b200.so (sm_100a), instructions per kernel\n")
    print("`cuobjdump -sass` mnemonic counts. DMMA = FP64 tensor-core MMA (mma.sync m8n8k4; tcgen05 has no FP64 kind);")
    print("UTMALDG = cp.async.bulk.tensor (TMA tensor-map tile load); UBLKCP = cp.async.bulk (1-D TMA copy); SYNCS =")
    print("mbarrier operations; LDGSTS = cp.async; UCGABAR = cluster barrier.\n")
    print("| kernel | " + " | ".join(WANT) + " | total |")
    print("|---|" + "---:|" * (len(WANT) + 1))
    for name, c in sorted(counts.items()):
        print(f"| `{name}` | " + " | ".join(str(c.get(k, 0)) for k in WANT) + f" | {sum(c.values())} |")


if __name__ == "__main__":
    sys.exit(main())
