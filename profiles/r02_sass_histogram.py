"""Regenerate profiles/sass_histogram_r02.txt: per-kernel counts of the SASS opcodes that prove what the kernels are built
from (tcgen05.mma / ld / st / commit, bulk copies, cp.async, mbarrier ops, packed fp32x2 arithmetic, MUFU, spills).

    python -m gamd_b200.build && python profiles/r02_sass_histogram.py > profiles/sass_histogram_r02.txt
"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJS = ["mp_tc2cta.o", "mp_tc.o", "enc_tc.o", "node_tc.o", "neighbor.o", "model_wide.o", "model_fp32.o", "thermostat.o", "integrate.o"]
KEYS = ["UTCHMMA", "LDTM", "STTM", "UTCBAR", "UBLKCP", "UTMALDG", "LDGSTS", "SYNCS", "MUFU.EX2", "MUFU.RCP", "MUFU.TANH",
        "MUFU.RSQ", "FFMA2", "FADD2", "FMUL2", "FFMA", "F2FP", "LDG", "STG", "LDS", "STS", "SHFL", "ATOM", "RED", "ELECT",
        "CCTL", "MEMBAR", "ERRBAR", "LDL", "STL", "BAR.SYNC", "DFMA"]

for o in OBJS:
    txt = subprocess.run(["cuobjdump", "-sass", os.path.join(ROOT, "gamd_b200", "lib", o)], capture_output=True, text=True).stdout
    cur, hist = None, collections.OrderedDict()
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            cur = re.sub(r"\(.*\)$", "", re.sub(r"\(anonymous namespace\)::", "", cur))
            hist[cur] = collections.Counter()
            continue
        m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur:
            hist[cur]["_total"] += 1
            for k in KEYS:
                if m.group(1) == k or m.group(1).startswith(k + "."):
                    hist[cur][k] += 1
    print(f"## {o}")
    for fn, h in hist.items():
        if h["_total"] >= 200:
            print(f"{fn}: total={h['_total']} " + " ".join(f"{k}={h[k]}" for k in KEYS if h[k]))
    print()
