"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel launches, total ms, share.

usage: ncu_launch_summary.py launches.csv "<command the list was taken from>" > summary.txt"""
import collections
import csv
import re
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 14 and r[0].isdigit()]
tot = collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    name = re.sub(r"\(.*", "", r[4])
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"slode::", "", name)
    v = float(r[14].replace(",", ""))
    unit = r[13]
    ms = v / 1e6 if unit in ("ns", "nsecond") else (v / 1e3 if unit in ("us", "usecond") else v)
    tot[name][0] += 1
    tot[name][1] += ms
total = sum(v[1] for v in tot.values())
print("ncu --metrics gpu__time_duration.sum --clock-control none, %s" % (sys.argv[2] if len(sys.argv) > 2 else ""))
print("kernel | launches | total ms | share of device time (per-launch times under ncu are serialised and cold-cache)")
for k, (n, ms) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print("%-95s %6d %10.3f %6.1f%%" % (k[:95], n, ms, 100 * ms / total))
