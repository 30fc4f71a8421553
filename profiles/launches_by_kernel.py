#!/usr/bin/env python
"""ncu launch list (--metrics gpu__time_duration.sum --csv) -> per-kernel table.  Usage: launches_by_kernel.py launches.csv [steps_in_capture] [skip_first_n_steps_fraction]
Counts every launch of the capture; the share column is each kernel's part of the summed GPU time (cold-cache, serialised: shares, not absolutes)."""
import csv
import re
import sys
from collections import defaultdict

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
head = rows[0]
ik, iv = head.index("Kernel Name"), head.index("Metric Value")
tot, cnt = defaultdict(float), defaultdict(int)
for r in rows[1:]:
    name = re.sub(r"^ipcgpu::", "", r[ik])
    name = re.sub(r"\(.*$", "", name)
    name = re.sub(r"^void ", "", name)
    if "cub::" in name:
        name = "cub::" + re.sub(r"<.*$", "", name.split("cub::")[-1].split("::")[-1])
    tot[name] += float(r[iv].replace(",", "")) / 1e3
    cnt[name] += 1
s = sum(tot.values())
w = csv.writer(sys.stdout)
w.writerow(["kernel", "launches", "us_per_launch", "us_total", "share"])
for k in sorted(tot, key=lambda k: -tot[k]):
    w.writerow([k, cnt[k], round(tot[k] / cnt[k], 1), round(tot[k], 1), round(tot[k] / s, 4)])
