"""Per-kernel time and DRAM traffic from an ncu launch list taken with
`--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv`.
usage: python profiles/summarize_traffic.py gpurun_out/launches.csv [top_n_launches] > profiles/<round>_traffic.txt"""
import collections
import csv
import re
import sys

UNIT = {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main(path, top=40):
    with open(path) as fh:
        lines = [l for l in fh if not l.startswith("==")]
    launches = collections.OrderedDict()
    for r in csv.DictReader(lines):
        rec = launches.setdefault(r["ID"], {"name": re.sub(r"\(.*", "", r["Kernel Name"])[:70], "grid": r.get("Grid Size", "")})
        rec[r["Metric Name"]] = float(r["Metric Value"].replace(",", "")) * UNIT.get(r["Metric Unit"], 1.0)
    per = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
    for rec in launches.values():
        p = per[rec["name"]]
        p[0] += 1
        p[1] += rec.get("gpu__time_duration.sum", 0.0)
        p[2] += rec.get("dram__bytes_read.sum", 0.0)
        p[3] += rec.get("dram__bytes_write.sum", 0.0)
    total = sum(p[1] for p in per.values())
    print(f"# {path}: {len(launches)} launches, {1e3 * total:.2f} ms of kernel time (cold-cache, serialised: compare SHARES)")
    print(f"# {'ms':>9} {'share':>6} {'n':>5} {'read GB':>8} {'write GB':>8} {'DRAM GB/s':>9}  kernel")
    for k, p in sorted(per.items(), key=lambda kv: -kv[1][1]):
        print(f"{1e3 * p[1]:11.3f} {100 * p[1] / total:5.1f}% {p[0]:5d} {p[2] / 1e9:8.2f} {p[3] / 1e9:8.2f} {(p[2] + p[3]) / max(p[1], 1e-12) / 1e9:9.0f}  {k}")
    print(f"\n# the {top} longest launches")
    for rec in sorted(launches.values(), key=lambda r: -r.get("gpu__time_duration.sum", 0.0))[:top]:
        t = rec.get("gpu__time_duration.sum", 0.0)
        b = rec.get("dram__bytes_read.sum", 0.0) + rec.get("dram__bytes_write.sum", 0.0)
        print(f"{1e3 * t:9.3f} ms  {rec.get('dram__bytes_read.sum', 0) / 1e9:6.2f} + {rec.get('dram__bytes_write.sum', 0) / 1e9:6.2f} GB  {b / max(t, 1e-12) / 1e9:6.0f} GB/s  {rec['name']}  {rec['grid']}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
