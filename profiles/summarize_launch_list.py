"""ncu launch list (csv of `--metrics gpu__time_duration.sum`) -> per-kernel table: launches, total us, share, avg us.

    python profiles/summarize_launch_list.py gpurun_out/launches.csv "the command that produced it" > profiles/rNN_launch_list_summary.txt
"""
import collections
import csv
import re
import sys

path = sys.argv[1]
rows = []
with open(path, newline="") as f:
    lines = [l for l in f if l.startswith('"')]
rd = csv.reader(lines)
hdr = next(rd)
iname, imetric, ival, iunit = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot = collections.defaultdict(float)
cnt = collections.Counter()
for r in rd:
    if len(r) <= ival or r[imetric] != "gpu__time_duration.sum":
        continue
    v = float(r[ival].replace(",", ""))
    unit = r[iunit]
    us = v / 1e3 if unit in ("ns", "nsecond") else (v * 1e3 if unit in ("ms", "msecond") else v)
    name = re.sub(r"\(.*", "", r[iname])[:110]
    tot[name] += us
    cnt[name] += 1
total = sum(tot.values())
if len(sys.argv) > 2:
    print(sys.argv[2])
print()
for name, us in sorted(tot.items(), key=lambda kv: -kv[1]):
    print(f"{cnt[name]:5d} {us:10.1f} us {100 * us / total:5.1f}%  avg {us / cnt[name]:7.2f} us  {name}")
print(f"total {total:.1f} us over {sum(cnt.values())} launches")
