"""Per-kernel SASS evidence of the tensor-core kernels in libtib.so: counts of tcgen05 MMAs (UTCHMMA), TMEM loads / stores
(LDTM / STTM), bulk copies of the TMA unit (UBLKCP), local-memory traffic (LDL / STL) and the resource usage.
    python tools/sass_summary.py > profiles/r02_sass_summary.txt"""
import collections
import os
import re
import subprocess

LIB = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "thermodynamic_interpolation_b200", "libtib.so")
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
usage = {}
cur = None
for line in res.splitlines():
    m = re.search(r"Function (\S+):", line)
    if m:
        cur = m.group(1)
    elif cur and "REG:" in line:
        usage[cur] = line.strip()
        cur = None
counts = collections.OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    for op in ("UTCHMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "UTMALDG", "LDL", "STL", "SYNCS", "MUFU", "BAR.SYNC"):
        if re.search(r"\b" + re.escape(op), line):
            counts[cur][op] += 1
print("# cuobjdump -sass / -res-usage of thermodynamic_interpolation_b200/libtib.so (sm_100a), tensor-core kernels")
print("# UTCHMMA = tcgen05.mma, UTCBAR = tcgen05.commit, LDTM / STTM = tcgen05.ld / st, UBLKCP = cp.async.bulk (TMA unit, 1-D),")
print("# UTMALDG = tensor-map TMA (none: weights are pre-swizzled 1-D blobs), LDL / STL = local memory")
for fn, c in counts.items():
    if c["UTCHMMA"] == 0:
        continue
    name = subprocess.run(["c++filt", fn], capture_output=True, text=True).stdout.strip()
    print(f"{name}\n    " + "  ".join(f"{k}:{c[k]}" for k in ("UTCHMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "UTMALDG", "LDL", "STL", "MUFU")) +
          f"\n    {usage.get(fn, '')}")
