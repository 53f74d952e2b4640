#!/usr/bin/env python
"""Benchmark of the LocAtE G+D training step (BASELINE.json metric: images/sec at 128x128).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--res 128] [--batch B] [--depth D]

A "step" = one discriminator update + one generator update on one synthetic batch (SURVEY.md 8d:
main.py:142-172 with miniter = MINIBATCHES = DITERS = 1).  Data parallel: one process per GPU
(torchrun), per-GPU batch fixed (weak scaling), global norm statistics / penalty means / gradients
all-reduced over NCCL.  Prints ONE JSON line on rank 0.

Besides the contract keys the line carries
  parity_check        one step at THIS configuration on the tensor-core path against the same step on the fp32 kernels
                      from identical state (losses, relative gradient-norm error), asserted before anything is timed;
  roofline            the dominant kernel family, SURVEY.md 8d arithmetic (bf16 activation bytes, MAC x 2), plus the ten
                      launch shapes that cost the most time, each judged against ITS roof (tensor if AI >= ridge else HBM);
  attention           the self-attention block (norm -> 1x1 -> RootTanh -> 1x1 -> softmax over HW -> gate) forward +
                      backward at the shapes the step uses: % of tensor-core peak, its min(1, F/ridge) cap, and the
                      same block under PyTorch eager on the same GPU;
  gpu_eager_baseline  the reference arithmetic (oracle port, pure torch) under PyTorch eager / cuDNN on the same GPU,
                      fp32 (TF32 convs, torch's default) and bf16 autocast -- SURVEY.md 2c's "kernel to beat on the box";
  cpu_baseline        the same port on the host cores.

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

# conv+linear MAC x 2 per image for one step = 4*F_G + 11*F_D, default flags
STEP_GFLOP_PER_IMAGE = {32: 1.44, 64: 7.46, 128: 36.2, 256: 172.1}
# the unmodified reference under torch.autocast(bfloat16) against itself in fp64 (tests/golden/autocast_yardstick.txt):
# relative gradient-norm error D / G.  The tensor-core path must not be worse than the reference's own bf16 mode.
YARDSTICK = {32: 1.5e-2, 64: 2.7e-2, 128: 2.8e-2, 256: 3.5e-2}


def metric_name(res):
    return f"G+D train-step images/sec at {res}x{res}"


def read_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            d = json.load(fh)
        return dict(hbm_gbs=d["hbm_gbs"], tflops=d["bf16_tflops_sustained"], tflops_burst=d["bf16_tflops"],
                    source="measured (MEASURED_PEAKS.json: copy bandwidth, cuBLAS bf16 sustained)")
    return dict(hbm_gbs=6650.0, tflops=1400.0, tflops_burst=1590.0, source="fallback (B200_PROFILING.md)")


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


# ------------------------------------------------------------------------------------------------------------------
# baselines built on the oracle port (test infrastructure: timed as a baseline, never on the product path)
# ------------------------------------------------------------------------------------------------------------------
def oracle_step_rate(res, depth, batch, steps, warmup, threads, device="cpu", autocast=False):
    """images/s of the oracle (pure-torch restatement of the reference) on the host cores, or -- device="cuda" -- under
    PyTorch eager / cuDNN on the GPU (fp32 with torch's default TF32 convolutions, or bf16 autocast)."""
    import contextlib
    import torch
    from oracle import locate_oracle as O
    if device == "cpu":
        torch.set_num_threads(threads)
    cfg = O.OracleConfig(IMAGE_SIZE=res, DEPTH=depth)
    gs, noise = O.init_generator_state(cfg, seed=999)
    ds = O.init_discriminator_state(cfg, seed=1000)
    if device != "cpu":
        gs = O.load_state({k: v.detach().to(device) for k, v in gs.items()})
        ds = O.load_state({k: v.detach().to(device) for k, v in ds.items()})
        noise = noise.to(device)
    g_opt = O.Nadam(cfg.GLR, (cfg.BETA_1, cfg.BETA_2))
    d_opt = O.Nadam(cfg.DLR, (cfg.BETA_1, cfg.BETA_2))
    real, aug, z = (t.to(device) for t in O.synthetic_batch(cfg, batch))
    ctx = (lambda: torch.autocast("cuda", dtype=torch.bfloat16)) if autocast else contextlib.nullcontext

    def sync():
        if device != "cpu":
            torch.cuda.synchronize()
    for _ in range(warmup):
        with ctx():
            O.train_step(gs, ds, noise, real, aug, z, cfg, g_opt, d_opt)
    sync()
    t0 = time.perf_counter()
    for _ in range(steps):
        with ctx():
            O.train_step(gs, ds, noise, real, aug, z, cfg, g_opt, d_opt)
    sync()
    dt = time.perf_counter() - t0
    return batch * steps / dt, dt / steps


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    threads = os.cpu_count() or 1
    batch, res = args.ref_batch, args.res
    value, sec = oracle_step_rate(res, args.depth, batch, args.steps, args.warmup, threads)
    sample = (f"{args.steps} timed + {args.warmup} warm-up full G+D steps (Nadam included) of the CPU oracle at "
              f"{res}x{res}, batch {batch} per step, fp32, {threads} torch threads")
    line = {
        "impl": "reference", "metric": metric_name(res), "value": value, "unit": "images/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(res, args.depth), "resolution": res,
                   "per_step_batch": batch, "parallelism": "cpu"},
        "cpu_baseline": {"value": value, "unit": "images/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


def workload_name(res, depth):
    flags = "default flags (DEPTH=1, attention on)" if depth == 1 else f"DEPTH={depth} bottleneck stacks, attention on"
    return f"G+D train step {res}x{res}, {flags}, Nadam included"


# ------------------------------------------------------------------------------------------------------------------
# correctness gate at the bench configuration
# ------------------------------------------------------------------------------------------------------------------
def parity_check(L, trainer, gen, dis, g_opt, d_opt, real, aug, z, res):
    """One step on the tensor-core path and the same step on the fp32 kernels (lb_conv_gemm / lb_conv_wgrad, the path
    tests/test_gpu_parity.py pins to the reference's golden vectors at rtol 1e-4) from IDENTICAL state and inputs, at
    the batch the benchmark runs: losses within 2e-2, relative gradient-norm error within the bf16 yard-stick."""
    import torch

    def snapshot():
        return ({k: v.detach().clone() for k, v in gen.state_dict().items()}, {k: v.detach().clone() for k, v in dis.state_dict().items()})

    def restore(snap):
        with torch.no_grad():
            for model, sd in zip((gen, dis), snap):
                cur = model.state_dict()
                for k, v in sd.items():
                    cur[k].copy_(v)
        L.ops.invalidate_packs()

    snap = snapshot()
    uv_flags = [(p, p.requires_grad) for p in dis.parameters()]
    results = {}
    for mode in ("bf16", "fp32"):
        restore(snap)
        for p, flag in uv_flags:
            p.requires_grad_(flag)
        L.configure(PRECISION=mode)
        grabbed = {}
        orig = {}
        for tag, opt in (("d", d_opt), ("g", g_opt)):
            orig[tag] = opt.step

            def spy(closure=None, tag=tag, opt=opt):
                grabbed[tag] = [g.detach().clone() for g in opt.flat_grads]      # no parameter update: state stays put
            opt.step = spy
        try:
            d_out, g_out = trainer._eager_step(real, aug, z)
            torch.cuda.synchronize()
        finally:
            for tag, opt in (("d", d_opt), ("g", g_opt)):
                opt.step = orig[tag]
        results[mode] = (d_out.detach().cpu().tolist(), g_out.detach().cpu().tolist(), grabbed)
    L.configure(PRECISION="bf16")
    restore(snap)
    for p, flag in uv_flags:
        p.requires_grad_(flag)
    (d16, g16, gr16), (d32, g32, gr32) = results["bf16"], results["fp32"]
    out = {"batch": int(real.shape[0]), "resolution": res,
           "reference_path": "fp32 SIMT kernels of this library (golden-fixture tier, tests/test_gpu_parity.py)",
           "d_hinge": [d16[0], d32[0]], "penalty": [d16[1], d32[1]], "g_hinge": [g16[0], g32[0]]}
    ok = True
    for name, a, b in (("d_hinge", d16[0], d32[0]), ("g_hinge", g16[0], g32[0])):
        rel = abs(a - b) / max(abs(b), 1e-12)
        out[name + "_rel_err"] = rel
        ok = ok and rel <= 2e-2
    pen_err = abs(d16[1] - d32[1])
    out["penalty_abs_err"] = pen_err
    ok = ok and pen_err <= 0.15 * abs(d32[1]) + 2e-6
    limit = YARDSTICK.get(res, 3.5e-2)
    for tag in ("d", "g"):
        a = torch.cat([t.double().reshape(-1) for t in gr16[tag]])
        b = torch.cat([t.double().reshape(-1) for t in gr32[tag]])
        rel = ((a - b).norm() / b.norm().clamp_min(1e-30)).item()
        out[f"{tag}_grad_rel_err"] = rel
        ok = ok and rel <= limit
    out["grad_rel_err_limit"] = limit
    out["ok"] = bool(ok)
    return out


# ------------------------------------------------------------------------------------------------------------------
# the attention block in isolation (BASELINE.json: "attention kernel % of TC peak"; config 5)
# ------------------------------------------------------------------------------------------------------------------
def attention_shapes(res):
    """(F, H) of every self-attention instance of G and D at this resolution (block.py:28-29: even blocks, size >= 8)."""
    import locate_b200 as L
    from locate_b200 import models
    out = []
    gf, df = models.generator_feature_list(), models.discriminator_feature_list()
    size = 2
    for i in range(len(gf) - 1):
        size *= 2
        if size >= L.CFG.MIN_ATTENTION_SIZE and i % L.CFG.ATTENTION_EVERY_NTH_LAYER == 0:
            out.append(("G", gf[i + 1], size))
    size = res // 2
    for i in range(len(df) - 1):
        size //= 2
        if size >= L.CFG.MIN_ATTENTION_SIZE and i % L.CFG.ATTENTION_EVERY_NTH_LAYER == 0:
            out.append(("D", df[i + 1], size))
    return out


def attention_bench(res, batch, peaks, reps=5):
    """ResModule(identity, Norm(F, SelfAttention(F))) forward + backward (block.py:42-43, attention.py:40-54), timed with
    CUDA events, against the same block written with torch ops (the oracle's functions) on the same GPU."""
    import torch
    import locate_b200 as L
    from locate_b200 import layers
    from oracle import locate_oracle as O
    ridge = peaks["tflops"] * 1e12 / (peaks["hbm_gbs"] * 1e9)
    rows = []
    for net, feat, size in attention_shapes(res):
        torch.manual_seed(11)
        m = layers.ResModule(layers.identity, layers.Norm(feat, layers.SelfAttention(feat))).to("cuda")
        from locate_b200 import ops
        st_dtype = ops.store_dtype((batch, feat, size, size))          # activations arrive in the model's storage type
        x32 = torch.randn((batch, feat, size, size), device="cuda").contiguous(memory_format=torch.channels_last)
        g32 = torch.randn((batch, feat, size, size), device="cuda").contiguous(memory_format=torch.channels_last)
        x = x32.to(st_dtype).requires_grad_(True)
        g = g32.to(st_dtype)

        def mine():
            x.grad = None
            m(x).backward(g)
        st = O.load_state({k: v.detach().clone() for k, v in m.state_dict().items()})
        xe = x32.contiguous().requires_grad_(True)
        ge = g32.contiguous()

        def eager(autocast=False):
            xe.grad = None
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
                h = O.whole_tensor_norm(xe, st["layer_module.i_norm.weight"], st["layer_module.i_norm.bias"])
                h = O.self_attention(st, "layer_module.module.", h, 4)
                y = O.gate(xe, h, st["gamma"], True)
            y.backward(ge.to(y.dtype))

        def timed(fn):
            for _ in range(2):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                fn()
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / reps
        # the block as it runs inside the step: replayed from a CUDA graph (the step is one graph); eager launches of this
        # library go through ctypes + autograd per kernel, which would dominate a 1 ms block
        how = "cuda graph replay"
        try:
            for _ in range(2):
                mine()
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                mine()
            ms = timed(graph.replay)
            del graph
        except Exception as exc:                       # noqa: BLE001
            how = f"eager launches (capture failed: {type(exc).__name__})"
            ms = timed(mine)
        ms_eager = timed(eager)
        ms_autocast = timed(lambda: eager(True))
        pixels = batch * size * size
        flops = 12.0 * pixels * feat * feat            # two FxF GEMMs: forward 2 x 2MF^2, backward 2 x (dgrad + wgrad)
        tf = flops / (ms * 1e-3) / 1e12
        cap = min(1.0, feat / ridge)                   # fused pair: AI = F flop/B
        rows.append({"net": net, "features": feat, "hw": size * size, "batch": batch, "ms_fwd_bwd": ms,
                     "timed_as": how, "tflops": tf, "frac_of_tc_peak": tf / peaks["tflops"], "roofline_cap": cap,
                     "frac_of_cap": tf / peaks["tflops"] / cap,
                     "torch_eager_fp32_ms": ms_eager, "torch_eager_bf16_autocast_ms": ms_autocast,
                     "speedup_vs_eager_fp32": ms_eager / ms, "speedup_vs_eager_bf16": ms_autocast / ms})
        del m, x, g, xe, ge, st, x32, g32
        torch.cuda.empty_cache()
    return rows


# ------------------------------------------------------------------------------------------------------------------
def shape_rows(ksum_by_label, peaks, steps, top=10):
    ridge = peaks["tflops"] * 1e12 / (peaks["hbm_gbs"] * 1e9)
    rows = []
    for (fam, label), v in ksum_by_label.items():
        if v["ms"] <= 0 or not label:
            continue
        ai = v["flops"] / max(v["bytes"], 1.0)
        tf = v["flops"] / (v["ms"] * 1e-3) / 1e12
        gbs = v["bytes"] / (v["ms"] * 1e-3) / 1e9
        tensor = ai >= ridge
        rows.append({"kernel": fam, "shape": label, "launches_per_step": v["launches"] / steps, "ms_per_step": v["ms"] / steps,
                     "ai_flop_per_byte": ai, "bound": "tensor" if tensor else "hbm", "tflops": tf, "gbs": gbs,
                     "frac": tf / peaks["tflops"] if tensor else gbs / peaks["hbm_gbs"]})
    rows.sort(key=lambda r: -r["ms_per_step"])
    return rows[:top]


def run_cuda(args):
    import torch
    import locate_b200 as L
    from locate_b200 import dist, ops

    rank, world = dist.init_from_env()
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    res = args.res
    L.configure(IMAGE_SIZE=res, DEPTH=args.depth, PRECISION=args.precision)
    torch.manual_seed(999)                      # replicas start identical (same seed, same RNG stream as the reference)
    gen, g_opt = L.get_model(L.Generator(), L.CFG.GLR, dev)
    dis, d_opt = L.get_model(L.Discriminator(), L.CFG.DLR, dev)
    trainer = L.GanTrainer(gen, dis, g_opt, d_opt)

    b = args.batch
    g = torch.Generator().manual_seed(1234 + rank)     # per-rank synthetic shard
    real_h = torch.randn((b, 3, res, res), generator=g).clamp_(-1, 1).pin_memory()
    aug_h = (real_h + 0.05 * torch.randn((b, 3, res, res), generator=g)).clamp_(-1, 1).pin_memory()
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

    # ---- correctness gate at THIS configuration (before anything is timed); one rank, no collectives involved
    parity = None
    if world == 1 and args.precision == "bf16" and not args.no_parity:
        trainer.step(real, aug, z)                          # one real step first: D's u / v become trainable (main.py:172)
        try:
            parity = parity_check(L, trainer, gen, dis, g_opt, d_opt, real, aug, z, res)
        except Exception as exc:                            # noqa: BLE001  (e.g. out of memory on the fp32 leg: say so)
            L.configure(PRECISION="bf16")
            parity = {"ok": None, "error": f"{type(exc).__name__}: {exc}"[:300]}
        if parity.get("ok") is False:
            print(json.dumps({"error": "parity_check failed at the bench configuration", "parity_check": parity}), flush=True)
            return 3
        torch.cuda.empty_cache()
        torch.cuda.reset_peak_memory_stats()                # `hbm_peak_gib` describes the measured bf16 step, not the gate's fp32 leg

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
    ms_local = e0.elapsed_time(e1)
    ms_total = max_over_ranks(ms_local)
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
    klabel = ktimer.summary(by_label=True)
    for table in (ksum, klabel):
        for v in table.values():                           # normalise to the `steps` of the timed region
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
    hbm_peak = round(torch.cuda.max_memory_allocated() / 2 ** 30, 2)

    if rank != 0:
        return 0

    peaks = read_peaks()
    ridge = peaks["tflops"] * 1e12 / (peaks["hbm_gbs"] * 1e9)
    fam = max(ksum, key=lambda k: ksum[k]["ms"]) if ksum else None
    roofline = None
    if fam:
        k = ksum[fam]
        per_launch_ms = k["ms"] / k["launches"]
        tflops = k["flops"] / (k["ms"] * 1e-3) / 1e12
        gbs = k["bytes"] / (k["ms"] * 1e-3) / 1e9
        # flops = MAC x 2; bytes = bf16 activations in + out + bf16 weights (fp32 weight gradient for the
        # wgrad kernel); the family is tensor-bound when its aggregate intensity is above the ridge
        ai = k["flops"] / max(k["bytes"], 1.0)
        tensor_bound = ai >= ridge
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
                    "frac": tflops / peaks["tflops"] if tensor_bound else gbs / peaks["hbm_gbs"],
                    "traffic": traffic, "peak_source": peaks["source"],
                    "ai_flop_per_byte": ai, "ridge_flop_per_byte": ridge,
                    "algorithmic_flops_per_step": k["flops"] / args.steps, "algorithmic_bytes_per_step": k["bytes"] / args.steps,
                    "launches_per_step": k["launches"] / args.steps, "avg_launch_ms": per_launch_ms,
                    "share_of_step": k["ms"] / ms_local,
                    "families": {n: {"ms_per_step": v["ms"] / args.steps, "tflops": v["flops"] / (v["ms"] * 1e-3) / 1e12,
                                     "gbs": v["bytes"] / (v["ms"] * 1e-3) / 1e9,
                                     "launches_per_step": v["launches"] / args.steps} for n, v in ksum.items()},
                    "shapes": shape_rows(klabel, peaks, args.steps)}
    if args.shapes_file:
        with open(args.shapes_file, "w") as fh:
            for r in shape_rows(klabel, peaks, args.steps, top=10 ** 6):
                fh.write(f"{r['ms_per_step']:8.3f} ms  n={r['launches_per_step']:5.1f}  {r['kernel']:16s} {r['shape']:52s} "
                         f"{r['bound']:6s} {r['tflops']:7.1f} TF/s {r['gbs']:7.1f} GB/s  frac {r['frac']:.3f}\n")
    gflop = STEP_GFLOP_PER_IMAGE.get(res) if args.depth == 1 else None
    line = {
        "metric": metric_name(res), "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16" if L.CFG.PRECISION == "bf16" else "f32", "data": "synthetic",
        "config": {"workload": workload_name(res, args.depth),
                   "precision": "bf16 tcgen05 GEMM operands, fp32 accumulate / reductions / optimizer"
                   if L.CFG.PRECISION == "bf16" else "fp32 everywhere",
                   "resolution": res, "per_gpu_batch": b, "global_batch": b * world, "parallelism": f"dp{world}",
                   "l2": "per-step working set (saved activations, several GB) exceeds the 126 MB L2; no explicit flush",
                   "norm_statistics": "global batch (all-reduced)" if world > 1 else "single process",
                   "launch": graph_note},
        "clocks": clocks.summary(),
        "e2e": {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": launches,
        "hbm_peak_gib": hbm_peak,
        "step_tflops": value * gflop / 1e3 if gflop else None,
        "step_frac_of_tc_peak": value * gflop / 1e3 / (world * peaks["tflops"]) if gflop else None,
        "roofline": roofline,
        "parity_check": parity if parity is not None else "skipped (runs on the single-GPU fp32-vs-bf16 configuration only)",
        "losses": {"d_hinge": float(losses[0][0]), "penalty": float(losses[0][1]), "g_hinge": float(losses[1][0])},
    }
    if world == 1:
        trainer.release_graph()
        del trainer
        torch.cuda.empty_cache()
        if not args.no_attention:
            line["attention"] = attention_bench(res, min(b, args.attention_batch), peaks)
        if not args.no_gpu_eager:
            eb = min(b, args.eager_batch)
            try:
                v32, s32 = oracle_step_rate(res, args.depth, eb, 3, 2, 1, device="cuda")
                v16, s16 = oracle_step_rate(res, args.depth, eb, 3, 2, 1, device="cuda", autocast=True)
                line["gpu_eager_baseline"] = {
                    "what": "the reference arithmetic (oracle port, pure torch ops) under PyTorch eager / cuDNN / cuBLAS on this GPU",
                    "batch": eb, "fp32_tf32conv_images_per_s": v32, "fp32_ms_per_step": s32 * 1e3,
                    "bf16_autocast_images_per_s": v16, "bf16_autocast_ms_per_step": s16 * 1e3,
                    "speedup_vs_fp32": value / v32, "speedup_vs_bf16_autocast": value / v16}
            except Exception as exc:                       # noqa: BLE001
                line["gpu_eager_baseline"] = {"error": f"{type(exc).__name__}: {exc}"[:200]}
            torch.cuda.empty_cache()
        if not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            cb, cs = args.ref_batch, 2
            v, sec = oracle_step_rate(res, args.depth, cb, cs, 1, threads)
            line["cpu_baseline"] = {"value": v, "unit": "images/s", "cores": threads, "kind": "port",
                                    "sample": f"{cs} timed + 1 warm-up full G+D steps of the CPU oracle at {res}x{res}, batch {cb}, "
                                              f"fp32, {threads} torch threads ({sec:.2f} s/step)"}
    print(json.dumps(line), flush=True)
    return 0


def main():
    # NCCL prints its banner / INFO log on stdout; rank 0 must print exactly one JSON line there, so the log goes to stderr
    # (it stays visible: the driver reads the communicator's rank count from it)
    if os.environ.get("NCCL_DEBUG") and not os.environ.get("NCCL_DEBUG_FILE"):
        os.environ["NCCL_DEBUG_FILE"] = "/dev/stderr"
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--res", type=int, default=128, choices=[32, 64, 128, 256], help="IMAGE_SIZE (BASELINE.json's metric is quoted at 128)")
    ap.add_argument("--depth", type=int, default=1, help="DEPTH (config.py:68); BASELINE config 4 uses 3 at 256x256")
    ap.add_argument("--batch", type=int, default=512, help="per-GPU batch (weak scaling: fixed per GPU)")
    ap.add_argument("--ref-batch", type=int, default=4, help="batch of the CPU reference sample")
    ap.add_argument("--eager-batch", type=int, default=64, help="batch of the PyTorch-eager-on-GPU baseline")
    ap.add_argument("--attention-batch", type=int, default=512, help="batch of the isolated attention-block measurement (capped by --batch)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-eager", action="store_true")
    ap.add_argument("--no-attention", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the fp32-vs-bf16 correctness gate")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel eagerly instead of replaying a CUDA graph")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--shapes-file", default=None, help="write every GEMM-class launch shape of a step with its time / roofline here")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_cuda(args)


if __name__ == "__main__":
    sys.exit(main())
