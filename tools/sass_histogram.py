#!/usr/bin/env python3
"""Opcode histogram per kernel of the shipped library, from `cuobjdump -sass` (no GPU needed).

    python tools/sass_histogram.py [metacoag_b200/libmcg_b200.so] > profiles/rNN_sass_histogram.txt

Per kernel: instruction count and the opcodes that say what hardware path it takes -- FP64
(DFMA / DMMA), async copies (LDGSTS = cp.async, UBLKCP / UTMALDG = TMA bulk copies), mbarrier
(SYNCS), tensor-core generations (HMMA / IMMA = mma.sync, UTCHMMA / UTCQMMA = tcgen05.mma, LDTM /
STTM = tensor memory), shared-memory and global accesses.
"""
import collections
import re
import subprocess
import sys

WATCH = ("DFMA", "DMMA", "DADD", "DMUL", "MUFU", "LDGSTS", "UBLKCP", "UTMALDG", "UTMASTG", "SYNCS",
         "UTCHMMA", "UTCQMMA", "UTCIMMA", "UTCBAR", "LDTM", "STTM", "UTCCP", "HMMA", "IMMA", "LDS", "STS",
         "LDG", "STG", "RED", "ATOM", "ATOMS", "ATOMG", "SHFL", "BAR", "LDSM", "FENCE", "ELECT")


def main():
    lib = sys.argv[1] if len(sys.argv) > 1 else "metacoag_b200/libmcg_b200.so"
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = kernels.setdefault(m.group(1), collections.Counter())
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m and cur is not None:
            cur[m.group(1)] += 1
    demangle = subprocess.run(["cu++filt"] + list(kernels), capture_output=True, text=True).stdout.splitlines()
    total = collections.Counter()
    print("# %s: opcode histogram per kernel (cuobjdump -sass)" % lib)
    for (name, hist), nice in zip(kernels.items(), demangle):
        nice = re.sub(r"\(.*", "", nice)
        picked = ["%s %d" % (op, sum(v for k, v in hist.items() if k == op or k.startswith(op + "_")))
                  for op in WATCH if any(k == op or k.startswith(op + "_") for k in hist)]
        print("%-60s %6d instr | %s" % (nice[:60], sum(hist.values()), ", ".join(picked)))
        total.update(hist)
    print("\n# whole library")
    for op in WATCH:
        n = sum(v for k, v in total.items() if k == op or k.startswith(op + "_"))
        if n:
            print("%-10s %d" % (op, n))


if __name__ == "__main__":
    main()
