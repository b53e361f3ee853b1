"""Summarise an ncu launch list (`--metrics gpu__time_duration.sum --csv`) of ONE training step: total
serialised kernel time and the share of every kernel.
usage: python tools/launch_summary.py profiles/r2_step_launches_b64.csv > profiles/r2_step_launches_b64_summary.txt"""
import collections
import csv
import re
import sys


def main(path):
    rows = list(csv.reader(open(path, errors="replace")))
    hdr = None
    acc = collections.OrderedDict()
    for r in rows:
        if len(r) > 5 and r[0] == "ID":
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            d = dict(zip(hdr, r))
            if d.get("Metric Name") != "gpu__time_duration.sum":
                continue
            name = re.sub(r"^(void )?sisr::(<unnamed>|\(anonymous namespace\))::", "", d["Kernel Name"])
            name = re.sub(r"\(.*", "", name)
            unit = d.get("Metric Unit", "ns")
            v = float(d["Metric Value"].replace(",", ""))
            v *= {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3}.get(unit, 1e-3)
            a = acc.setdefault(name, [0.0, 0])
            a[0] += v
            a[1] += 1
    total = sum(a[0] for a in acc.values())
    n = sum(a[1] for a in acc.values())
    print(f"total {total / 1e3:.3f} ms over {n} launches")
    for name, (us, cnt) in sorted(acc.items(), key=lambda kv: -kv[1][0]):
        print(f"  {us / 1e3:6.3f} ms {100 * us / total:5.1f}% {cnt:5d} x {us / cnt:8.1f} us  {name}")


if __name__ == "__main__":
    main(sys.argv[1])
