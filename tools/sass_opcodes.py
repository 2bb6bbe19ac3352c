#!/usr/bin/env python3
"""Per-kernel SASS opcode histogram of libecseg_b200.so: the Blackwell-native proof (B200_PROFILING.md, "What proves a
Blackwell-native kernel").  Runs on the CPU box:

    python tools/sass_opcodes.py [lib.so] > profiles/r02_sass_opcodes.txt

    tcgen05.mma -> UTC*MMA (UTCHMMA for kind::f16)     tcgen05.ld / .st -> LDTM / STTM       tcgen05.commit -> UTCBAR
    cp.async.bulk.tensor load / store -> UTMALDG / UTMASTG      mma.sync (legacy) -> HMMA     wgmma -> HGMMA (must be 0)
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEY = ["UTCHMMA", "UTCQMMA", "UTCBAR", "UTCATOM", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "UBLKCP", "SYNCS", "HMMA", "HGMMA",
       "VIMNMX3", "REDUX", "ATOM", "RED"]


def main():
    lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "ecseg_b200", "libecseg_b200.so")
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    fn = None
    per = collections.OrderedDict()
    arch = set()
    for ln in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", ln)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = re.sub(r"\(anonymous namespace\)::", "", name)
            name = re.sub(r"\(.*\)$", "", name).replace("void ecseg::", "")
            fn = name
            per[fn] = collections.Counter()
            continue
        m = re.match(r"\s*arch = (\S+)", ln)
        if m:
            arch.add(m.group(1))
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", ln)
        if m and fn:
            per[fn][m.group(1)] += 1
    print(f"# cuobjdump -sass {os.path.relpath(lib, ROOT)}   arch: {', '.join(sorted(arch))}")
    print(f"# {len(per)} kernels; columns = instruction counts per kernel (static SASS)")
    cols = [k for k in KEY if any(c[k] for c in per.values())]
    print("kernel".ljust(66) + "".join(k.rjust(9) for k in cols) + "   total")
    tot = collections.Counter()
    for fn, c in per.items():
        if not any(c[k] for k in cols):
            continue
        print(fn[:65].ljust(66) + "".join(str(c[k]).rjust(9) for k in cols) + str(sum(c.values())).rjust(8))
        for k in cols:
            tot[k] += c[k]
    print("ALL KERNELS".ljust(66) + "".join(str(tot[k]).rjust(9) for k in cols))
    rest = [fn for fn, c in per.items() if not any(c[k] for k in cols)]
    print(f"# {len(rest)} kernels without any of these opcodes (byte / integer kernels): " + ", ".join(sorted(set(rest)))[:1500])
    assert tot["HGMMA"] == 0 and tot["HMMA"] == 0, "legacy tensor-core path present"


if __name__ == "__main__":
    main()
