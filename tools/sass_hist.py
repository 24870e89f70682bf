#!/usr/bin/env python
"""Opcode histogram per kernel of the product library (cuobjdump -sass), for profiles/.

    python tools/sass_hist.py [--lib x265-mod-by-patman_b200/lib/libx265b200.so] [--match REGEX] [--top N] > profiles/rN_sass_histograms.md

Instructions are counted statically (one per SASS line), predicates stripped, modifiers dropped (IMAD.WIDE -> IMAD); the Blackwell
mnemonics B200_PROFILING.md names (UTCIMMA / UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG = TMA bulk tensor
copies, UTCBAR = tcgen05.commit) are listed in a column of their own so that their presence or absence is visible at a glance.
"""
import argparse
import collections
import os
import re
import subprocess

HERE = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BLACKWELL = ("UTCIMMA", "UTCHMMA", "UTCQMMA", "UTCOMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCBAR", "UTCCP", "SYNCS", "UBLKCP")


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.split("\n")
    return dict(zip(names, out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--lib", default=os.path.join(HERE, "x265-mod-by-patman_b200", "lib", "libx265b200.so"))
    ap.add_argument("--match", default=".")
    ap.add_argument("--top", type=int, default=14)
    a = ap.parse_args()
    txt = subprocess.run(["cuobjdump", "-sass", a.lib], capture_output=True, text=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    arch = set()
    for line in txt.split("\n"):
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = kernels.setdefault(m.group(1), collections.Counter())
            continue
        m = re.search(r"arch = (sm_\w+)", line)
        if m:
            arch.add(m.group(1))
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", line)
        if m and cur is not None:
            ins = re.sub(r"^@!?U?P\w+\s+", "", m.group(1).strip())
            op = ins.split()[0].split(".")[0] if ins else ""
            if op:
                cur[op] += 1
    names = demangle(list(kernels))
    sel = [(names[k], c) for k, c in kernels.items() if re.search(a.match, names[k])]
    total = collections.Counter()
    for _, c in sel:
        total.update(c)
    print("# SASS opcode histograms, %s (%s), %d kernels" % (os.path.basename(a.lib), ", ".join(sorted(arch)), len(sel)))
    print()
    print("Whole library: " + ", ".join("%s %d" % kv for kv in total.most_common(24)))
    print()
    print("Blackwell-specific mnemonics in the library: " + (", ".join("%s %d" % (k, total[k]) for k in BLACKWELL if total[k]) or "none"))
    print("Legacy tensor / packed-integer mnemonics: " + ", ".join("%s %d" % (k, total[k]) for k in ("IMMA", "HMMA", "IDP", "VABSDIFF", "VABSDIFF4", "VIMNMX", "VIADDMNMX", "I2IP") if total[k]))
    print()
    print("| kernel | instructions | Blackwell mnemonics | top opcodes |")
    print("|---|---|---|---|")
    for name, c in sorted(sel, key=lambda x: x[0]):
        short = re.sub(r"\(.*", "", name).replace("void ", "").replace("b200::", "")
        bw = ", ".join("%s %d" % (k, c[k]) for k in BLACKWELL if c[k]) or "-"
        print("| `%s` | %d | %s | %s |" % (short[:110], sum(c.values()), bw, ", ".join("%s %d" % kv for kv in c.most_common(a.top))))


if __name__ == "__main__":
    main()
