#!/usr/bin/env python
"""Per-kernel summary of an `ncu --metrics ... --csv` launch list: median of every metric per kernel name.
usage: launch_times.py <launches.csv> [skip_first_n_launches]"""
import collections, csv, statistics, sys
rows = list(csv.reader(open(sys.argv[1])))
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
i = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
hdr = rows[i]
acc = collections.OrderedDict()
for r in rows[i + 1:]:
    if len(r) != len(hdr) or int(r[0]) < skip:
        continue
    name = r[4].split("(")[0].replace("void ", "").replace("rdpn::", "")
    acc.setdefault(name, collections.OrderedDict()).setdefault(r[-3], []).append(float(r[-1].replace(",", "")))
for name, ms in acc.items():
    print(name, " ".join("%s=%.4g" % (k.split("__")[-1].split(".")[0], statistics.median(v)) for k, v in ms.items()), "n=%d" % len(next(iter(ms.values()))))
