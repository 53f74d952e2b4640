#!/usr/bin/env python
"""Benchmark of the LocAtE G+D training step at 128x128 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--batch B]

A "step" = one discriminator update + one generator update on one synthetic batch (SURVEY.md 8d:
main.py:142-172 with miniter = MINIBATCHES = DITERS = 1).  Data parallel: one process per GPU
(torchrun), per-GPU batch fixed (weak scaling), global norm statistics / penalty means / gradients
all-reduced over NCCL.  Prints ONE JSON line on rank 0.

--impl reference: the reference's CPU arithmetic (the oracle port of it -- the Python reference itself
cannot travel to the GPU box) timed on the host cores for the same metric and config.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

RES = 128
METRIC = "G+D train-step images/sec at 128x128"
# conv+linear MAC x 2 per image for one step = 4*F_G + 11*F_D
STEP_GFLOP_PER_IMAGE = 36.2


def read_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            d = json.load(fh)
        return dict(hbm_gbs=d["hbm_gbs"], tflops=d["bf16_tflops_sustained"], source="measured (MEASURED_PEAKS.json, sustained)")
    return dict(hbm_gbs=6650.0, tflops=1400.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def __exit__(self, *exc):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except subprocess.TimeoutExpired:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def oracle_step_rate(batch, steps, warmup, threads):
    """images/s of the CPU oracle (restatement of the reference) on the host cores."""
    import torch
    from oracle import locate_oracle as O
    torch.set_num_threads(threads)
    cfg = O.OracleConfig(IMAGE_SIZE=RES)
    gs, noise = O.init_generator_state(cfg, seed=999)
    ds = O.init_discriminator_state(cfg, seed=1000)
    g_opt = O.Nadam(cfg.GLR, (cfg.BETA_1, cfg.BETA_2))
    d_opt = O.Nadam(cfg.DLR, (cfg.BETA_1, cfg.BETA_2))
    real, aug, z = O.synthetic_batch(cfg, batch)
    for _ in range(warmup):
        O.train_step(gs, ds, noise, real, aug, z, cfg, g_opt, d_opt)
    t0 = time.perf_counter()
    for _ in range(steps):
        O.train_step(gs, ds, noise, real, aug, z, cfg, g_opt, d_opt)
    dt = time.perf_counter() - t0
    return batch * steps / dt, dt / steps


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    threads = os.cpu_count() or 1
    batch = args.ref_batch
    value, sec = oracle_step_rate(batch, args.steps, args.warmup, threads)
    sample = (f"{args.steps} timed + {args.warmup} warm-up full G+D steps (Nadam included) of the CPU oracle at "
              f"{RES}x{RES}, batch {batch} per step, fp32, {threads} torch threads")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "images/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"G+D train step {RES}x{RES}, default flags (DEPTH=1, attention on)", "resolution": RES,
                   "per_step_batch": batch, "parallelism": "cpu"},
        "cpu_baseline": {"value": value, "unit": "images/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


def run_cuda(args):
    import torch
    import locate_b200 as L
    from locate_b200 import dist, ops

    rank, world = dist.init_from_env()
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    L.configure(IMAGE_SIZE=RES, PRECISION=args.precision)
    torch.manual_seed(999)                      # replicas start identical (same seed, same RNG stream as the reference)
    gen, g_opt = L.get_model(L.Generator(), L.CFG.GLR, dev)
    dis, d_opt = L.get_model(L.Discriminator(), L.CFG.DLR, dev)
    trainer = L.GanTrainer(gen, dis, g_opt, d_opt)

    b = args.batch
    g = torch.Generator().manual_seed(1234 + rank)     # per-rank synthetic shard
    real_h = torch.randn((b, 3, RES, RES), generator=g).clamp_(-1, 1).pin_memory()
    aug_h = (real_h + 0.05 * torch.randn((b, 3, RES, RES), generator=g)).clamp_(-1, 1).pin_memory()
    z_h = torch.randn((b, L.CFG.INPUT_VECTOR_Z), generator=g).pin_memory()
    real, aug, z = real_h.to(dev), aug_h.to(dev), z_h.to(dev)
    # inputs stay NCHW like the reference's loaders deliver them; the NCHW -> channels-last kernel is part of the step

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
            return float(t.item())
        return ms

    # warm-up: W eager steps; then (default) the step is captured into ONE CUDA graph
    graph_note = "eager launches"
    if args.no_graph:
        for _ in range(args.warmup):
            trainer.step(real, aug, z)
    else:
        try:
            trainer.capture(real, aug, z, warmup=args.warmup)
            trainer.step(real, aug, z)                     # one untimed replay
            graph_note = "whole step replayed as one CUDA graph"
        except Exception as exc:                           # noqa: BLE001  (report, then measure eagerly)
            trainer.release_graph()
            graph_note = f"eager launches (graph capture failed: {type(exc).__name__}: {exc})"[:300]
            for _ in range(args.warmup):
                trainer.step(real, aug, z)
    barrier()

    # ---- timed region 1: device-resident inputs -> `value`
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clocks:
        barrier()
        e0.record()
        for _ in range(args.steps):
            trainer.step(real, aug, z)
        e1.record()
        barrier()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    ms_step = ms_total / args.steps
    value = world * b * args.steps / (ms_total * 1e-3)

    # ---- per-kernel CUDA events for the roofline: the same kernels launched eagerly (events cannot be read
    # back from inside a graph), two extra untimed steps; also counts the launches of one step
    L.reset_launch_count()
    with ops.KernelTimer() as ktimer:
        for _ in range(2):
            trainer._eager_step(real, aug, z)
    barrier()
    launches = L.launch_count() // 2 * args.steps          # kernels of this library executed per step x timed steps
    ksum = ktimer.summary()
    for v in ksum.values():                                # normalise to the `steps` of the timed region
        for key in ("launches", "flops", "bytes", "ms"):
            v[key] = v[key] / 2 * args.steps

    # ---- timed region 2: end to end through the public API with HOST buffers (H2D + D2H inside): every step's batch is
    # copied from pinned host memory (prefetch() of batch i+1 runs on a side stream while step i computes -- the first
    # copy is inside the timed region too) and every step's losses are read back to the host
    barrier()
    e0.record()
    trainer.prefetch(real_h, aug_h, z_h)
    for i in range(args.steps):
        d_out, g_out = trainer.step_prefetched()
        if i + 1 < args.steps:
            trainer.prefetch(real_h, aug_h, z_h)
        losses = (d_out.cpu(), g_out.cpu())             # device -> host read of the step's result
    e1.record()
    barrier()
    ms_e2e = max_over_ranks(e0.elapsed_time(e1))
    e2e_value = world * b * args.steps / (ms_e2e * 1e-3)
    h2d = (real_h.numel() + aug_h.numel() + z_h.numel()) * 4
    d2h = (3 + 1) * 4

    if rank != 0:
        return 0

    peaks = read_peaks()
    fam = max(ksum, key=lambda k: ksum[k]["ms"]) if ksum else None
    roofline = None
    if fam:
        k = ksum[fam]
        per_launch_ms = k["ms"] / k["launches"]
        tflops = k["flops"] / (k["ms"] * 1e-3) / 1e12
        gbs = k["bytes"] / (k["ms"] * 1e-3) / 1e9
        # a kernel family that mixes tensor-bound (C >= 192 transposed convs) and HBM-bound launches
        # (1x1 / C <= 96) is judged against whichever roof it sits closer to; both views are reported.
        frac_tc, frac_hbm = tflops / peaks["tflops"], gbs / peaks["hbm_gbs"]
        tensor_bound = frac_tc >= frac_hbm
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if os.path.exists(tpath):
            with open(tpath) as fh:
                traffic = json.load(fh).get(fam)
        names = {"conv_gemm": "k_conv_gemm (fp32 SIMT fwd+dgrad gather-GEMM)", "conv_wgrad": "k_conv_wgrad (fp32 SIMT)",
                 "conv_tc": "k_conv_tc2 + k_conv_tc (tcgen05 implicit GEMM fwd+dgrad, persistent / split-K)",
                 "wgrad_tc": "k_wgrad_tc (tcgen05 wgrad)", "conv_small": "k_conv_small (direct fp32, tiny channel counts)",
                 "conv_small_wgrad": "k_conv_small_wgrad"}
        roofline = {"kernel": names.get(fam, fam),
                    "bound": "tensor" if tensor_bound else "hbm",
                    "achieved": tflops if tensor_bound else gbs,
                    "peak": peaks["tflops"] if tensor_bound else peaks["hbm_gbs"],
                    "unit": "TFLOP/s" if tensor_bound else "GB/s",
                    "frac": frac_tc if tensor_bound else frac_hbm,
                    "traffic": traffic, "peak_source": peaks["source"],
                    "tensor_view": {"achieved": tflops, "peak": peaks["tflops"], "unit": "TFLOP/s", "frac": frac_tc},
                    "hbm_view": {"achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": frac_hbm,
                                 "algorithmic_bytes_per_step": k["bytes"] / args.steps},
                    "launches_per_step": k["launches"] / args.steps, "avg_launch_ms": per_launch_ms,
                    "share_of_step": k["ms"] / (ms_total if world == 1 else e0.elapsed_time(e1)),
                    "families": {n: {"ms_per_step": v["ms"] / args.steps, "tflops": v["flops"] / (v["ms"] * 1e-3) / 1e12,
                                     "gbs": v["bytes"] / (v["ms"] * 1e-3) / 1e9,
                                     "launches_per_step": v["launches"] / args.steps} for n, v in ksum.items()}}
    line = {
        "metric": METRIC, "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16" if L.CFG.PRECISION == "bf16" else "f32", "data": "synthetic",
        "config": {"workload": f"G+D train step {RES}x{RES}, default flags (DEPTH=1, attention on), Nadam included",
                   "precision": "bf16 tcgen05 GEMM operands, fp32 accumulate / activations / reductions / optimizer"
                   if L.CFG.PRECISION == "bf16" else "fp32 everywhere",
                   "resolution": RES, "per_gpu_batch": b, "global_batch": b * world, "parallelism": f"dp{world}",
                   "l2": "per-step working set (saved activations, several GB) exceeds the 126 MB L2; no explicit flush",
                   "norm_statistics": "global batch (all-reduced)" if world > 1 else "single process",
                   "launch": graph_note},
        "clocks": clocks.summary(),
        "e2e": {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": launches,
        "hbm_peak_gib": round(torch.cuda.max_memory_allocated() / 2 ** 30, 2),
        "step_tflops": value * STEP_GFLOP_PER_IMAGE / 1e3,
        "roofline": roofline,
        "losses": {"d_hinge": float(losses[0][0]), "penalty": float(losses[0][1]), "g_hinge": float(losses[1][0])},
    }
    if world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        cb, cs = args.ref_batch, 2
        v, sec = oracle_step_rate(cb, cs, 1, threads)
        line["cpu_baseline"] = {"value": v, "unit": "images/s", "cores": threads, "kind": "port",
                                "sample": f"{cs} timed + 1 warm-up full G+D steps of the CPU oracle at {RES}x{RES}, batch {cb}, "
                                          f"fp32, {threads} torch threads ({sec:.2f} s/step)"}
    print(json.dumps(line), flush=True)
    return 0


def main():
    os.environ.pop("NCCL_DEBUG", None)         # NCCL prints its version banner on stdout at any debug level: rank 0 prints exactly one JSON line
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--batch", type=int, default=512, help="per-GPU batch (weak scaling: fixed per GPU); ~34 GB of HBM at 512")
    ap.add_argument("--ref-batch", type=int, default=4, help="batch of the CPU reference sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel eagerly instead of replaying a CUDA graph")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_cuda(args)


if __name__ == "__main__":
    sys.exit(main())
