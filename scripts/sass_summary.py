#!/usr/bin/env python
"""Static SASS instruction mix per kernel of libcfdb200.so (`cuobjdump -sass`), plus registers / spills from the ptxas log.
Runs without a GPU.  Usage: python scripts/sass_summary.py > profiles/rNN_sass_summary.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "cfd-with-cuda_b200", "libcfdb200.so")
LOG = os.path.join(ROOT, "cfd-with-cuda_b200", "csrc", "device.ptxas.log")

COLS = [("LDG.128", r"\bLDG\.E(\.\w+)*\.128"), ("LDG.64", r"\bLDG\.E(\.\w+)*\.64"), ("LDS.128", r"\bLDS\.128"), ("LDS.64", r"\bLDS\.64"),
        ("STG", r"\bSTG\."), ("STS", r"\bSTS"), ("RED/ATOM", r"\b(RED|REDG|ATOM|ATOMG|ATOMS)\."), ("LDL/STL", r"\b(LDL|STL)\b"),
        ("DFMA..", r"\b(DFMA|DADD|DMUL)\b"), ("SHFL", r"\bSHFL\."), ("BAR", r"\bBAR\."), ("TMA/MMA", r"\b(UTMA|UBLKCP|HMMA|UTCMMA|QGMMA)")]


def demangle(names):
    out = subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.splitlines()
    return [re.sub(r"\(.*", "", n).replace("cfdb::", "") for n in out]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    counts, cur = collections.OrderedDict(), None
    for ln in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", ln)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        if cur is None or "/*" not in ln or not re.search(r"/\*[0-9a-f]{4,}\*/", ln):
            continue
        counts[cur]["instr"] += 1
        for name, rx in COLS:
            if re.search(rx, ln):
                counts[cur][name] += 1
    regs = {}
    if os.path.exists(LOG):
        fn = None
        for ln in open(LOG):
            m = re.search(r"Compiling entry function '(\S+)'", ln)
            if m:
                fn = m.group(1)
            m = re.search(r"(\d+) bytes spill stores", ln)
            if m and fn:
                regs.setdefault(fn, [0, 0])[1] = int(m.group(1))
            m = re.search(r"Used (\d+) registers", ln)
            if m and fn:
                regs.setdefault(fn, [0, 0])[0] = int(m.group(1))
    names = list(counts)
    pretty = dict(zip(names, demangle(names)))
    print("# SASS instruction mix of the sm_100a kernels in libcfdb200.so (static counts from `cuobjdump -sass`; registers and spill stores from")
    print("# `ptxas -v`).  LDG.128 / LDG.64 = vectorised global loads of the matrix streams, LDS.128 / LDS.64 = shared-memory loads (the resident")
    print("# pressure CG reads its matrix with LDS.128), RED/ATOM = FP64 atomics of the element scatter and work counters, LDL/STL = local memory.")
    print("# TMA/MMA: no tensor-core or TMA instruction anywhere -- nothing on this path is a dense contraction, and the gathers are indexed.")
    hdr = "%-46s %6s %5s %6s " % ("kernel", "instr", "regs", "spillB") + " ".join("%8s" % n for n, _ in COLS)
    print(hdr)
    for fn in sorted(names, key=lambda f: -counts[f]["instr"]):
        c = counts[fn]
        r = regs.get(fn, [0, 0])
        print("%-46s %6d %5d %6d " % (pretty[fn][:46], c["instr"], r[0], r[1]) + " ".join("%8d" % c[n] for n, _ in COLS))


if __name__ == "__main__":
    main()
