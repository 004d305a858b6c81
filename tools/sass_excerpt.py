#!/usr/bin/env python
"""SASS evidence (read here, no GPU): per kernel of libs2ag_b200.so the counts of the Blackwell-specific mnemonics
(UTCHMMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UBLKCP = cp.async.bulk (1-D TMA), UTMALDG = tensor-map TMA,
UTCBAR = tcgen05.commit, SYNCS = mbarrier, LDGSTS = cp.async, FFMA2 = packed fp32 FMA, UCGABAR/ CGA = cluster barrier,
LDG.E.256-class wide loads) plus the first occurrence of each as an excerpt.
usage: python tools/sass_excerpt.py > profiles/r02_sass_summary.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "speech2affective_gestures_b200", "libs2ag_b200.so")
KEYS = ["UTCHMMA", "LDTM", "STTM", "UBLKCP", "UTMALDG", "UTMASTG", "UTCBAR", "SYNCS", "LDGSTS", "FFMA2", "UCGABAR",
        "LDG.E.ENL2.256", "LDG.E.128", "STG.E.ENL2.256", "RED.E", "ATOMG", "MUFU", "HMMA", "BAR.SYNC"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    kern, counts, first, n_instr = None, {}, {}, collections.Counter()
    arch = set()
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            kern = re.sub(r"\(.*", "", kern)[:110]
            counts.setdefault(kern, collections.Counter())
            first.setdefault(kern, {})
            continue
        m = re.match(r"\s*arch = (\S+)", line)
        if m:
            arch.add(m.group(1))
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(.*?);", line)
        if kern and m:
            ins = m.group(1)
            n_instr[kern] += 1
            for k in KEYS:
                if re.search(r"(^|\s)" + re.escape(k), ins):
                    counts[kern][k] += 1
                    first[kern].setdefault(k, ins.strip())
    print("# cuobjdump -sass speech2affective_gestures_b200/libs2ag_b200.so ; arch:", ", ".join(sorted(arch)))
    tot = collections.Counter()
    for k in counts.values():
        tot.update(k)
    print("# totals:", ", ".join("%s %d" % (k, tot[k]) for k in KEYS if tot[k]))
    print()
    for kern in sorted(counts, key=lambda k: -counts[k]["UTCHMMA"]):
        c = counts[kern]
        if not any(c[k] for k in KEYS[:11]):
            continue
        print("%s  [%d SASS instructions]" % (kern, n_instr[kern]))
        print("    " + ", ".join("%s %d" % (k, c[k]) for k in KEYS if c[k]))
        for k in ("UTCHMMA", "LDTM", "UBLKCP", "UTCBAR", "LDGSTS", "FFMA2", "UCGABAR"):
            if k in first[kern]:
                print("      %-8s e.g.  %s" % (k, first[kern][k][:120]))
    print()
    print("# kernels without tensor-core / TMA / mbarrier instructions (SIMT, HBM- or latency-bound):")
    rest = [k for k in counts if not any(counts[k][x] for x in KEYS[:11])]
    for kern in sorted(rest):
        c = counts[kern]
        print("  %-100s %s" % (kern, ", ".join("%s %d" % (k, c[k]) for k in KEYS if c[k])))


if __name__ == "__main__":
    main()
