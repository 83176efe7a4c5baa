#!/bin/bash
# ncu launch list (per-launch gpu__time_duration) of a python command; usage: tools/launch_list.sh out.csv python ...
out=$1; shift
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file "$out" "$@" > /dev/null 2>&1
python - "$out" <<'PY'
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]; ik = hdr.index("Kernel Name"); iv = hdr.index("Metric Value"); iu = hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[1:]:
    v = float(r[iv].replace(",", "")); u = r[iu]
    v = v / 1e3 if u in ("nsecond", "ns") else (v if u in ("usecond", "us") else v * 1e3)
    k = r[ik].split("(")[0]
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{t/1e3:10.3f} ms {100*t/tot:5.1f}%  {n:5d} launches  {t/n:10.1f} us/launch  {k}")
print(f"{tot/1e3:10.3f} ms total")
PY
