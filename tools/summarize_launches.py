"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel (and grid size).
usage: python tools/summarize_launches.py launches.csv [--grid]"""
import collections, csv, re, sys

def main():
    path = sys.argv[1]
    by_grid = "--grid" in sys.argv
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    agg = collections.defaultdict(lambda: [0, 0.0])
    tot = 0.0
    for row in csv.DictReader(lines):
        k = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "").replace("sisr::<unnamed>::", "")
        if by_grid:
            k = f"{k} grid={row['Grid Size']}"
        v = float(row["Metric Value"])
        agg[k][0] += 1; agg[k][1] += v; tot += v
    print(f"total {tot/1e6:.3f} ms over {sum(a[0] for a in agg.values())} launches")
    for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"{t/1e6:8.3f} ms {100*t/tot:5.1f}% {n:5d} x {t/n/1e3:8.1f} us  {k}")

if __name__ == "__main__":
    main()
