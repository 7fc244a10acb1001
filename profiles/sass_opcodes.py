#!/usr/bin/env python
"""Blackwell-native proof a reader can check: per-kernel counts of the sm_100a tensor-core / TMA / TMEM SASS opcodes in
etpnav_b200/libetpnav_b200.so (cuobjdump -sass).  UTCHMMA = tcgen05.mma (``.2CTA`` = cta_group::2), UTMALDG / UTMASTG =
TMA tile load / store (cp.async.bulk.tensor), LDTM / STTM = tcgen05.ld / st (TMEM), UTCBAR = tcgen05.commit,
SYNCS = mbarrier ops.  HMMA (mma.sync) must be absent.

    python profiles/sass_opcodes.py > profiles/r02_sass_opcodes.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "etpnav_b200", "libetpnav_b200.so")
OPS = ["UTCHMMA.2CTA", "UTCHMMA", "UTMALDG", "UTMASTG", "LDTM", "STTM", "UTCBAR", "UTCATOMSWS", "SYNCS", "HMMA", "MUFU.EX2", "RED.E"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    counts, cur = collections.OrderedDict(), None
    for ln in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", ln)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
        if not m:
            continue
        op = m.group(1)
        for o in OPS:
            if op.startswith(o):
                # "UTCHMMA" also prefixes "UTCHMMA.2CTA": count the 2CTA form separately
                if o == "UTCHMMA" and ".2CTA" in op:
                    continue
                counts[cur][o] += 1
                break
    demangle = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
    print("# " + " ".join(sys.argv))
    print("# SASS opcode counts per kernel of etpnav_b200/libetpnav_b200.so (sm_100a); only kernels with at least one listed opcode")
    print("\t".join(["kernel"] + OPS))
    tot = collections.Counter()
    for (k, c), name in zip(counts.items(), demangle):
        if not any(c[o] for o in OPS):
            continue
        name = re.sub(r"\(anonymous namespace\)::", "", name)
        name = re.sub(r"\(.*", "", name)[:110]
        print("\t".join([name] + [str(c[o]) for o in OPS]))
        tot.update(c)
    print("\t".join(["TOTAL"] + [str(tot[o]) for o in OPS]))


if __name__ == "__main__":
    main()
