#!/usr/bin/env python3
"""Per-kernel counts of the SASS mnemonics that prove a Blackwell-native kernel (B200_PROFILING.md): UTC*MMA = tcgen05.mma,
LDTM / STTM = tcgen05.ld / st, UTCBAR = tcgen05.commit, UBLKCP = cp.async.bulk (1-D TMA), SYNCS = mbarrier, ACQBULK = griddepcontrol.wait
(programmatic dependent launch), plus HMMA (legacy mma.sync: must be absent).   python profiles/sass_counts.py > profiles/r2_sass_counts.txt"""
import os
import re
import subprocess
import sys
from collections import OrderedDict

so = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "diffusion_edf_b200", "libdedf.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
MN = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTCBAR", "UTCATOMSWS", "UBLKCP", "UTMALDG", "SYNCS", "ACQBULK", "HMMA", "FFMA", "REDUX", "MUFU"]
cur, counts = None, OrderedDict()
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        counts.setdefault(cur, dict.fromkeys(MN, 0))
        continue
    if cur:
        for k in MN:
            if re.search(r"\b" + k + r"\b", line) or (k in ("UTCHMMA", "UTCQMMA", "SYNCS", "MUFU", "UTCBAR", "LDTM", "STTM", "ACQBULK", "UBLKCP") and k in line):
                counts[cur][k] += 1
arch = subprocess.run(["cuobjdump", "-lelf", so], capture_output=True, text=True).stdout
print("# cuobjdump -sass diffusion_edf_b200/libdedf.so (sm_100a only: %s)" % ", ".join(sorted(set(re.findall(r"sm_\d+a?", arch)))))
print("# mnemonic counts per kernel; kernels without any tensor / TMA / mbarrier instruction are listed with FFMA only")
print(f"{'kernel':58s} " + " ".join(f"{k:>8s}" for k in MN))
for name, c in counts.items():
    print(f"{name[:58]:58s} " + " ".join(f"{c[k]:8d}" for k in MN))
