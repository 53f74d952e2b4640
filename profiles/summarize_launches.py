"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel name.
usage: python profiles/summarize_launches.py gpurun_out/launches.csv > profiles/<round>_launches.txt"""
import collections
import csv
import re
import sys


def main(path):
    with open(path) as fh:
        lines = [l for l in fh if not l.startswith("==")]
    tot = collections.defaultdict(lambda: [0, 0.0])
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", r["Kernel Name"])[:80]
        v = float(r["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3}.get(r["Metric Unit"], 1.0)
        tot[name][0] += 1
        tot[name][1] += v
    s = sum(v[1] for v in tot.values())
    print(f"# {path}: {sum(v[0] for v in tot.values())} launches, {s:.2f} ms of kernel time (cold-cache, serialised)")
    print(f"# {'ms':>10} {'share':>6} {'launches':>8}  kernel")
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        print(f"{v[1]:12.3f} {100 * v[1] / s:5.1f}% {v[0]:8d}  {k}")


if __name__ == "__main__":
    main(sys.argv[1])
