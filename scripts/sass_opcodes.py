#!/usr/bin/env python
"""Per-kernel counts of the Blackwell-specific SASS opcodes in libmvip_nerf.so (cuobjdump -sass), as evidence that the kernels
are tcgen05 / TMEM / bulk-TMA code:  UTC*MMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UBLKCP = cp.async.bulk (1-D TMA),
UTMALDG / UTMASTG = cp.async.bulk.tensor, UTCBAR = tcgen05.commit, SYNCS = mbarrier ops, HMMA = legacy mma.sync (should be 0).

    python scripts/sass_opcodes.py > profiles/r02_sass_opcodes.md
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "mvip_nerf_b200", "libmvip_nerf.so")
OPS = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UBLKCP", "UTMALDG", "UTMASTG", "UTCBAR", "SYNCS", "HMMA", "LDGSTS", "REDUX", "SHFL", "MUFU"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    counts, order, cur, total = {}, [], None, collections.Counter()
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            order.append(cur)
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and cur:
            op = m.group(1)
            total[cur] += 1
            for o in OPS:
                if op.startswith(o):
                    counts[cur][o] += 1
    demangle = subprocess.run(["cu++filt"] + order, capture_output=True, text=True).stdout.splitlines() if order else []
    names = dict(zip(order, demangle)) if len(demangle) == len(order) else {o: o for o in order}
    print("# r02 - Blackwell-specific SASS opcodes per kernel of `mvip_nerf_b200/libmvip_nerf.so`\n")
    print("`python scripts/sass_opcodes.py` (cuobjdump -sass, sm_100a).  UTCHMMA = `tcgen05.mma` kind::f16, LDTM / STTM = `tcgen05.ld` / `st`, "
          "UBLKCP = `cp.async.bulk` (1-D TMA, global<->shared), UTCBAR = `tcgen05.commit`, SYNCS = mbarrier ops; HMMA (legacy `mma.sync`) must be 0.\n")
    print("| kernel | SASS instrs | " + " | ".join(OPS) + " |")
    print("|---|---:|" + "---:|" * len(OPS))
    for f in order:
        nm = names[f]
        if nm.endswith(")"):                      # drop the parameter list (the last top-level parenthesis group)
            depth = 0
            for i in range(len(nm) - 1, -1, -1):
                depth += (nm[i] == ")") - (nm[i] == "(")
                if depth == 0:
                    nm = nm[:i]
                    break
        nm = nm.replace("(anonymous namespace)::", "").replace("<unnamed>::", "").replace("void ", "")
        print("| `%s` | %d | " % (nm[:70], total[f]) + " | ".join(str(counts[f][o]) for o in OPS) + " |")


if __name__ == "__main__":
    sys.exit(main())
