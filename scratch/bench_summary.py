"""Print the headline fields of a bench.py JSON line.  usage: python scratch/bench_summary.py file.json [n_shapes]"""
import json
import sys

r = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
n = int(sys.argv[2]) if len(sys.argv) > 2 else 0
print("value", round(r["value"], 1), r["unit"], "| ms/step", round(r["ms_per_step"], 2), "| e2e", round(r["e2e"]["value"], 1),
      "| launches", r.get("gpu_launches"), "| hbm GiB", r.get("hbm_peak_gib"), "| clocks", r.get("clocks"))
pc = r.get("parity_check")
if isinstance(pc, dict):
    print("parity", pc.get("ok"), "d_grad", pc.get("d_grad_rel_err"), "g_grad", pc.get("g_grad_rel_err"))
rf = r["roofline"]
print("roofline frac", round(rf["frac"], 4), rf["bound"], "achieved", round(rf["achieved"], 1), rf["unit"], "share", round(rf.get("share_of_step", 0), 3))
for k, v in rf.get("families", {}).items():
    print(f"  {k:18s} {v['ms_per_step']:7.2f} ms/step  {v['tflops']:7.1f} TF/s  {v['gbs']:7.1f} GB/s  x{v['launches_per_step']:.0f}")
for s in rf.get("shapes", [])[:n]:
    print(f"  {s['kernel']:16s} {s['shape']:52s} x{s['launches_per_step']:.0f} {s['ms_per_step']:6.3f} ms {s['bound']:6s} frac {s['frac']:.3f}")
if "scaling_detail" in r:
    print(r["scaling_detail"])
