#!/usr/bin/env python
"""Writes profiles/sass_tma_excerpt.txt: SASS evidence of the TMA / mbarrier path in the library as built
(cuobjdump -sass astre_b200/libastre_gpu.so; runs on the CPU box).  Per kernel: the counts of UBLKCP (cp.async.bulk =
TMA bulk copy), SYNCS (mbarrier), ATOMG / RED, LDG / STG, ACQBULK (griddepcontrol.wait = programmatic dependent launch),
CCTL.PF (prefetch); then the instructions themselves for the headline kernel k_frame_fused<false,false,true>."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "astre_b200", "libastre_gpu.so")
OUT = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "sass_tma_excerpt.txt")
HEADLINE = "k_frame_fusedILb0ELb0ELb1E"

sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
kernels, name = collections.OrderedDict(), None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        name = m.group(1)
        kernels[name] = []
    elif name and re.match(r"\s*/\*[0-9a-f]{4,}\*/", line):
        kernels[name].append(line.rstrip())

WANT = [("UBLKCP", r"\bUBLKCP"), ("SYNCS", r"\bSYNCS\.(ARRIVE|PHASECHK|EXCH)"), ("ATOMG", r"\bATOMG"), ("RED", r"\bRED\."),
        ("LDG", r"\bLDG"), ("STG", r"\bSTG"), ("ACQBULK", r"\bACQBULK"), ("PREFETCH", r"\bCCTL\S*\.PF")]
with open(OUT, "w") as f:
    f.write("# SASS evidence of the TMA / mbarrier path (tools/sass_excerpt.py: cuobjdump -sass astre_b200/libastre_gpu.so, sm_100a build of this commit)\n")
    f.write("# per kernel: count of UBLKCP (cp.async.bulk = TMA bulk copy), SYNCS (mbarrier ops), ATOMG/RED, LDG/STG, ACQBULK (PDL wait), "
            "PREFETCH (CCTL .PF); then the lines themselves for k_frame_fused<0,0,1>\n\n")
    for k, lines in kernels.items():
        counts = [(n, sum(1 for l in lines if re.search(rx, l))) for n, rx in WANT]
        f.write(f"{k[:90]:90s} " + " ".join(f"{n}={c}" for n, c in counts if c) + "\n")
    f.write("\n# k_frame_fused<false,false,true> (the headline kernel): the TMA, mbarrier and atomic instructions\n")
    for k, lines in kernels.items():
        if HEADLINE in k:
            for l in lines:
                if re.search(r"UBLKCP|SYNCS|ATOMG|\bRED\.|ACQBULK", l):
                    f.write(l.strip() + "\n")
print(f"wrote {OUT}: {len(kernels)} kernels")
