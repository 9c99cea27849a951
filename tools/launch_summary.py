"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) of one train step: per kernel
family, launches and summed duration.   python tools/launch_summary.py CSV [first_launch last_launch]"""
import collections
import csv
import re
import sys


def main():
    rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
    hdr = rows[0]
    ki, vi, gi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size")
    data = rows[1:]
    if len(sys.argv) > 3:
        data = data[int(sys.argv[2]):int(sys.argv[3])]
    agg = collections.OrderedDict()
    for r in data:
        name = re.sub(r"\(.*", "", r[ki]).replace("void ", "").replace("sdumc::", "")
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += float(r[vi]) / 1e3
    tot = sum(v[1] for v in agg.values())
    print(f"{len(data)} launches, {tot / 1e3:.3f} ms summed (serialised, cold-cache)")
    for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{us:9.1f} us  {100 * us / tot:5.1f}%  x{n:<4d} {k}")


if __name__ == "__main__":
    main()
