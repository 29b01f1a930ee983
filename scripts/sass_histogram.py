#!/usr/bin/env python
"""SASS opcode histogram of every kernel in libbreedgym_b200.so (cuobjdump -sass), the evidence that the contraction
kernels really are tcgen05 / TMEM / TMA code:

    UTCIMMA / UTCQMMA ...  tcgen05.mma          STTM / LDTM   tcgen05.st / tcgen05.ld (tensor memory)
    UTMALDG                cp.async.bulk.tensor  UBLKCP        cp.async.bulk (1-D TMA)
    LDGSTS                 cp.async              UTCBAR        tcgen05.commit
    SYNCS                  mbarrier ops          VOTE / SHF / LOP3 / IMAD  the Threefry / ballot path

    python scripts/sass_histogram.py > profiles/r02_sass_histogram.txt
"""
import collections
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
LIB = ROOT / "breedgym_b200" / "libbreedgym_b200.so"
KEY = ["UTCIMMA", "UTCHMMA", "UTCQMMA", "STTM", "LDTM", "UTMALDG", "UTMASTG", "UBLKCP", "UBLKPF", "LDGSTS", "UTCBAR", "SYNCS",
       "ATOMG", "RED", "VOTE", "SHF", "LOP3", "IMAD", "IADD3", "LDG", "STG", "LDS", "STS", "BAR"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True, check=True).stdout
    kernels, cur = collections.OrderedDict(), None
    archs = sorted(set(re.findall(r"arch = (sm_\w+)", out)))
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = re.sub(r"\(anonymous namespace\)::", "", name)
            cur = kernels.setdefault(name.split("(")[0], collections.Counter())
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m and cur is not None:
            cur[m.group(1)] += 1
    print(f"# {LIB.name}: cubin architectures {archs}; instruction counts per kernel (static SASS)")
    print(f"{'kernel':58s} {'total':>6s} " + " ".join(f"{k:>7s}" for k in KEY))
    for name, c in kernels.items():
        print(f"{name[:58]:58s} {sum(c.values()):6d} " + " ".join(f"{c.get(k, 0):7d}" for k in KEY))
    tot = collections.Counter()
    for c in kernels.values():
        tot.update(c)
    print(f"{'ALL':58s} {sum(tot.values()):6d} " + " ".join(f"{tot.get(k, 0):7d}" for k in KEY))


if __name__ == "__main__":
    sys.exit(main())
