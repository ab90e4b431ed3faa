#!/usr/bin/env python
"""bench.py — ResNet-50 bs256 bf16 eval-forward throughput on B200 (BASELINE.json's metric), one JSON line.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--model resnet50] [--batch 256] [--dtype bf16]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...     (N > 1)
    python bench.py --impl reference ...      # the UNMODIFIED reference (baseline/_ref) on the host cores
    python bench.py --scaling strong ...      # global batch fixed (256 -> 256/N images per GPU) instead of weak scaling

A "step" is one eval forward of the named network over one synthetic batch.
  value      images/s with the fp32 NCHW batch resident in HBM (CUDA events, W >= 3 warm-ups, K timed steps, max over ranks)
  sustained  the same loop run for >= 2 s (the K-step region of a 3 ms step is a burst-clock sample)
  e2e        images/s through the public module call with the batch in pinned HOST memory: H2D of the images and D2H of the
             logits inside the timed region, double-buffered.  `e2e` uses the uint8 image contract of the ingest kernel
             (decoded images, normalisation fused; a quarter of the bytes); `e2e_f32` is the reference's fp32 NCHW contract.
  parity     the timed step's own output against the CPU oracle on a subsample of its images
  roofline   the most expensive kernel launch of the step (per-op CUDA events; algorithmic FLOPs / bytes per SURVEY 8d)
  roofline_step  whole step against the sum of per-kernel bounds
  configs    (N = 1, default model only) the other four BASELINE.json configs measured the same way, briefly
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# model -> (default batch, H, W, tier).  Tiers: BASELINE.json pins bf16 for ResNet-50 and fp32 for ResNet-18; MobileNetV2 and
# SE-ResNeXt run in the fp16 storage tier (same kernels / MMA rate; the tier in which MobileNetV2 meets 2e-2 + top-1)
CONFIGS = {
    "resnet50": (256, 224, 224, "bf16"), "resnet18": (8, 224, 224, "fp32"), "mobilenetv2_w1": (256, 224, 224, "fp16"),
    "seresnext50_32x4d": (256, 224, 224, "fp16"), "deeplabv3_resnetd50b_voc": (16, 480, 480, "bf16"),
    "efficientnet_b0": (256, 224, 224, "fp16"),   # SURVEY 8(f) rank 1 (not BASELINE configs: measured to the same bar)
    "mobilenetv3_large_w1": (256, 224, 224, "fp16"),
}
BASELINE_CONFIGS = ["resnet18", "resnet50", "mobilenetv2_w1", "seresnext50_32x4d", "deeplabv3_resnetd50b_voc"]
TOL = {"bf16": 2e-2, "fp16": 2e-2, "fp32": 1e-4}   # north star: max|d|/max|ref|, identical top-1
MEAN, STD = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "tf_burst": d["bf16_tflops"], "tf_sustained": d["bf16_tflops_sustained"],
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "tf_burst": 1590.0, "tf_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 50 ms while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.lines, self.proc, self.start = index, [], None, 0

    def ready(self, timeout: float = 5.0):
        """Block until nvidia-smi has printed its first sample: its start-up (NVML attach, ~0.1-1 s, during which the driver
        can stall kernel launches for tens of ms) is then over."""
        t0 = time.time()
        while self.proc is not None and not self.lines and time.time() - t0 < timeout:
            time.sleep(0.02)
        time.sleep(0.1)

    def mark(self):
        """The timed region starts now: samples taken before (NVML start-up, warm-up steps) are not summarised.  The
        sampler is started BEFORE the warm-up so that nvidia-smi's own initialisation never overlaps a timed step."""
        self.start = len(self.lines)

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def __exit__(self, *exc):
        if self.proc is not None:
            time.sleep(0.25)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except subprocess.TimeoutExpired:
                self.proc.kill()

    def summary(self, start: int | None = None, end: int | None = None) -> dict:
        sm, smax, pw, reasons = [], 0, 0.0, set()
        for ln in self.lines[self.start if start is None else start:end]:
            f = [s.strip() for s in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax = max(smax, float(f[2])); pw = max(pw, float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": smax or None, "power_w_max": pw or None,
                "reasons": sorted(reasons), "samples": len(sm)}


def build_net(model: str, H: int, W: int, get_model=None):
    """The benchmarked weights: the reference's own random init under torch.manual_seed(0), same on every rank."""
    import torch
    if get_model is None:
        import pytorchcv_b200 as P
        get_model = P.get_model
    torch.manual_seed(0)
    kw = {"in_size": (H, W)} if model.startswith("deeplab") else {}
    return get_model(model, pretrained=False, **kw).eval()


def reference_get_model():
    """get_model of the UNMODIFIED reference package: baseline/_ref (pip --target install of /root/reference, travels to
    the GPU box) or /root/reference itself; None when neither is importable."""
    for path in (os.path.join(ROOT, "baseline", "_ref"), "/root/reference"):
        if os.path.isdir(os.path.join(path, "pytorchcv")):
            sys.path.insert(0, path)
            try:
                from pytorchcv.model_provider import get_model
                return get_model, path
            except Exception:  # noqa: BLE001
                sys.path.remove(path)
    return None, None


def first(y):
    return y[0] if isinstance(y, (tuple, list)) else y


def bind_to_gpu_numa_node(local: int) -> dict:
    """Pin this rank's host threads (and, by first touch, its pinned staging buffers) to the NUMA node of its GPU."""
    info = {"node": None, "cpus": None}
    try:
        import torch
        bus = torch.cuda.get_device_properties(local).pci_bus_id
        dom = torch.cuda.get_device_properties(local).pci_domain_id
        dev = torch.cuda.get_device_properties(local).pci_device_id
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev:02x}.0/numa_node"
        node = int(open(path).read().strip())
        info["node"] = node
        if node >= 0:
            cpus = set()
            for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
            allowed = os.sched_getaffinity(0) & cpus
            if allowed:
                os.sched_setaffinity(0, allowed)
                info["cpus"] = len(allowed)
    except Exception:  # noqa: BLE001
        pass
    return info


# ---------------------------------------------------------------------------------------------------------------------
# reference arm: the reference's own CPU forward on the host cores
# ---------------------------------------------------------------------------------------------------------------------
def run_reference(a) -> None:
    import torch
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    batch, H, W = a.batch, a.h, a.w
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    get_model, where = reference_get_model()
    x = torch.randn(batch, 3, H, W, generator=torch.Generator().manual_seed(1234))
    if get_model is not None:
        net = build_net(a.model, H, W, get_model)
        kind, how = "reference", f"unmodified pytorchcv from {os.path.relpath(where, ROOT) if where.startswith(ROOT) else where}"

        def fwd(t):
            with torch.no_grad():
                return net(t)
    else:
        from oracle import oracle_forward
        net = build_net(a.model, H, W)
        kind, how = "port", "oracle/ref_forward.py (reference package not importable here)"

        def fwd(t):
            return oracle_forward(net, t)
    # the full batch of the named config per step; warm-up / steps as asked, bounded by a wall-clock budget
    t0 = time.perf_counter()
    fwd(x)
    t1 = time.perf_counter() - t0
    warm = max(1, min(a.warmup, int(a.ref_budget_s * 0.2 / max(t1, 1e-3))))
    steps = max(1, min(a.steps, int(a.ref_budget_s * 0.8 / max(t1, 1e-3))))
    for _ in range(warm - 1):
        fwd(x)
    t0 = time.perf_counter()
    for _ in range(steps):
        fwd(x)
    dt = (time.perf_counter() - t0) / steps
    v = batch / dt
    line = {"impl": "reference", "metric": f"{a.model} bs{batch} {a.dtype} eval inference images/sec", "value": round(v, 2),
            "unit": "images/s", "n_gpus": a.gpus, "steps": steps, "warmup": warm, "ms_per_step": round(dt * 1e3, 2),
            "higher_is_better": True, "scaling": a.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_string(a.model, H, W, batch),
                       "global_batch": batch * (a.gpus if a.scaling == "weak" else 1), "sample_batch": batch,
                       "parallelism": f"reference CPU path on rank 0's host cores (torch fp32, all threads; {world} rank(s) launched)"},
            "cpu_baseline": {"value": round(v, 2), "unit": "images/s", "cores": cores, "kind": kind,
                             "sample": f"{steps} timed forwards of the full batch of {batch} images after {warm} warm-up(s), "
                                       f"torch {torch.__version__} CPU, {torch.get_num_threads()} threads; {how}"},
            "e2e": {"value": round(v, 2), "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def workload_string(model, H, W, batch):
    return (f"{model} eval forward, {H}x{W}, batch {batch} per GPU, fp32 NCHW in -> fp32 logits out "
            f"(random-init weights, torch.manual_seed(0))")


# ---------------------------------------------------------------------------------------------------------------------
# b200 arm
# ---------------------------------------------------------------------------------------------------------------------
def parity_check(model, net_cpu, x_cpu, y_dev, dtype, max_images=8):
    """The timed step's own output against the CPU oracle on a subsample of its images (eval-mode images are independent)."""
    import torch
    from oracle import oracle_forward
    N = x_cpu.shape[0]
    k = min(N, 2 if x_cpu.shape[-1] > 256 else max_images)
    idx = torch.linspace(0, N - 1, k).round().long().unique()
    want = oracle_forward(net_cpu, x_cpu[idx])
    wants = want if isinstance(want, (tuple, list)) else (want,)
    gots = y_dev if isinstance(y_dev, (tuple, list)) else (y_dev,)
    rel = 0.0
    for g, w in zip(gots, wants):
        g = g[idx.to(g.device)].float().cpu()
        rel = max(rel, float((g - w).abs().max() / (w.abs().max() + 1e-30)))
    g0, w0 = gots[0][idx.to(gots[0].device)].float().cpu(), wants[0]
    if w0.dim() == 2:
        top1 = bool(torch.equal(g0.argmax(1), w0.argmax(1)))
        agree = None
    else:
        agree = float((g0.argmax(1) == w0.argmax(1)).float().mean())
        top1 = agree >= 0.97
    tol = TOL[dtype]
    out = {"rel_err": float(f"{rel:.3e}"), "top1_equal": top1, "tolerance": tol, "within_tolerance": bool(rel <= tol),
           "images": [int(i) for i in idx], "vs": "oracle/ref_forward.py (torch CPU fp32) on the same weights and images"}
    if agree is not None:
        out["pixel_argmax_agreement"] = round(agree, 4)
    return out


def measure(a, model, N, H, W, dtype, K, Wm, rank, world, local, dev, pk, full: bool):
    """One config on this rank's GPU.  `full`: the headline config (sustained run, per-kernel roofline, CPU baseline)."""
    import torch
    import torch.distributed as dist
    import pytorchcv_b200 as P
    from pytorchcv_b200 import _lib, parallel

    net_cpu = build_net(model, H, W)
    net = build_net(model, H, W).to(dev)
    fast = P.accelerate(net, dtype=dtype, graph=bool(a.graph), check_weights=False)
    x_cpu = torch.randn(N, 3, H, W, generator=torch.Generator().manual_seed(1234 + rank))
    x = x_cpu.to(dev)
    runner = parallel.ShardedInference(fast, rank, world, exchange=a.exchange)

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local])
        torch.cuda.synchronize(dev)

    def timed(fn, steps, clk=None):
        import gc
        gc.disable()   # a generation-2 collection over the module trees built above costs tens of ms: not part of a step
        barrier()
        if clk is not None:
            clk.mark()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        barrier()
        gc.enable()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item() / steps

    # ---------------- value: inputs resident in HBM ----------------
    y = runner(x)
    cm = fast.compiled(x)
    keep = {}

    def step(i):
        keep["y"] = runner(x)
    import gc
    gc.collect()
    runner(x)
    probe_ms = timed(step, 3)   # rough step time: sizes the sustained loop and the load-up phases (same on every rank)
    # The GPU must be under continuous load before a short timed region: after any idle gap (compile, allocation, NVML
    # start-up) the clocks need ~100 ms of work to come back, and a 20-step region (58 ms) that starts too early was
    # measured anywhere between 45 k and 89 k img/s.  So: the >= 2 s sustained loop runs FIRST, then - without a gap - the
    # W warm-up steps and the K timed steps of the contract.
    sus = None
    with ClockSampler(local) as clk:
        clk.ready()                # nvidia-smi / NVML start-up happens here, not inside a timed region
        if full:
            n_sus = max(K, int(a.sustain_s * 1e3 / max(probe_ms, 1e-3)) + 1)
            ms_sus = timed(step, n_sus, clk)
            s0, s1 = clk.start, len(clk.lines)
            sus = {"value": round(N * world / (ms_sus * 1e-3), 1), "ms_per_step": round(ms_sus, 4), "steps": n_sus,
                   "seconds": round(ms_sus * n_sus / 1e3, 2), "clocks": clk.summary(s0, s1)}
        else:
            for _ in range(max(Wm, int(300.0 / max(probe_ms, 1e-3)))):   # ~0.3 s of load for the secondary configs
                runner(x)
        for _ in range(Wm):
            runner(x)
        barrier()
        l0 = _lib.launch_count()
        ms_step = timed(step, K, clk)
    launches = _lib.launch_count() - l0
    y = keep["y"]
    value = N * world / (ms_step * 1e-3)
    res = {"model": model, "batch": N, "dtype": dtype, "value": round(value, 1), "ms_per_step": round(ms_step, 4),
           "clocks": clk.summary(), "gpu_launches": int(launches)}
    if sus is not None:
        res["sustained"] = sus

    # ---------------- parity of the timed step's own output (this rank's shard of the gathered result) ----------------
    y_local = y
    if world > 1:
        sl = slice(rank * N, (rank + 1) * N)
        y_local = type(y)(t[sl] for t in y) if isinstance(y, (tuple, list)) else y[sl]
    if rank == 0:
        res["parity"] = parity_check(model, net_cpu, x_cpu, y_local, dtype)

    # ---------------- e2e: host batch -> H2D -> forward -> D2H result, double-buffered ----------------
    def e2e_run(make_host, fast_e, label):
        run_e = parallel.ShardedInference(fast_e, rank, world, exchange=a.exchange)
        xh = [make_host(i).pin_memory() for i in range(2)]
        xd = [torch.empty_like(xh[0], device=dev) for _ in range(2)]
        o = run_e(xd[0].copy_(xh[0]))
        outs = o if isinstance(o, (tuple, list)) else (o,)
        yh = [[torch.empty(t.shape, dtype=torch.float32).pin_memory() for t in outs] for _ in range(2)]
        comp, copy, back = torch.cuda.current_stream(dev), torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
        copied = [torch.cuda.Event() for _ in range(2)]
        consumed = [torch.cuda.Event() for _ in range(2)]
        done = torch.cuda.Event()

        def e2e_step(i):
            # three streams: images in (copy) | forward + exchange (comp) | results out (back); the D2H of step i overlaps the
            # forward of step i+1 (at N GPUs every rank reads back the GATHERED logits: N x 1 MB)
            b = i & 1
            with torch.cuda.stream(copy):
                copy.wait_event(consumed[b])
                xd[b].copy_(xh[b], non_blocking=True)
                copied[b].record(copy)
            comp.wait_event(copied[b])
            out = run_e(xd[b])
            consumed[b].record(comp)
            outs_i = out if isinstance(out, (tuple, list)) else (out,)
            with torch.cuda.stream(back):
                back.wait_event(consumed[b])
                for dst, src in zip(yh[b], outs_i):
                    dst.copy_(src, non_blocking=True)
                    src.record_stream(back)
                done.record(back)
            if i == last_step[0]:
                comp.wait_event(done)   # the timed region ends when the last result is on the host

        last_step = [-1]
        for i in range(max(Wm, int(300.0 / max(probe_ms, 1e-3)))):   # ~0.3 s of load: clocks are up when the timed region starts
            e2e_step(i)
        last_step[0] = K - 1
        ems = timed(e2e_step, K)
        return {"value": round(N * world / (ems * 1e-3), 1), "unit": "images/s",
                "h2d_bytes_per_step": xh[0].numel() * xh[0].element_size(),
                "d2h_bytes_per_step": sum(t.numel() * 4 for t in yh[0]), "ms_per_step": round(ems, 4),
                "input_contract": label, "pipelining": "2 pinned host + 2 device buffers; H2D, forward and D2H on three streams"}

    aff = ([1.0 / (255.0 * s) for s in STD], [-m / s for m, s in zip(MEAN, STD)])
    fast_u8 = P.accelerate(net, dtype=dtype, graph=bool(a.graph), check_weights=False, input_affine=aff)
    res["e2e"] = e2e_run(
        lambda i: torch.randint(0, 256, (N, 3, H, W), generator=torch.Generator().manual_seed(99 + rank + i), dtype=torch.uint8),
        fast_u8, "uint8 NCHW host images; (x/255 - mean)/std fused into the ingest kernel (pcv_stem_s2d_ingest_ex)")
    res["e2e_f32"] = e2e_run(
        lambda i: torch.randn(N, 3, H, W, generator=torch.Generator().manual_seed(99 + rank + i)),
        fast, "fp32 NCHW host tensor (the reference's forward signature)")
    res["exchange"] = runner.exchange_used if world > 1 else None

    if rank != 0:
        return res

    # ---------------- per-kernel roofline (rank 0): CUDA events around every op of the plan ----------------
    reps = 3
    rows = None
    for _ in range(reps):
        r = cm.profile()
        rows = r if rows is None else [(n, t0 + t1, f, b) for (n, t0, f, b), (_, t1, _, _) in zip(rows, r)]
    rows = [(n, t / reps, f, b) for n, t, f, b in rows]
    inside_step = K * ms_step > 2000.0
    F = (pk["tf_sustained"] if inside_step else pk["tf_burst"]) * 1e12
    BW = pk["hbm_gbs"] * 1e9
    table, sum_bound, sum_meas = [], 0.0, 0.0
    for n, t, f, b in rows:
        tb = max(f / F, b / BW) * 1e3
        table.append({"op": n, "ms": round(t, 4), "gflop": round(f / 1e9, 3), "mb": round(b / 1e6, 3),
                      "bound": "tensor" if f / F >= b / BW else "hbm", "t_bound_ms": round(tb, 4),
                      "frac": round(tb / t, 3) if t > 0 else None})
        sum_bound += tb
        sum_meas += t
    top = max(table, key=lambda r: r["ms"])
    if top["bound"] == "tensor":
        ach = top["gflop"] / top["ms"]  # GFLOP/ms == TFLOP/s
        roof = {"bound": "tensor", "achieved": round(ach, 1), "peak": round(F / 1e12, 1), "unit": "TFLOP/s",
                "frac": round(ach / (F / 1e12), 3)}
    else:
        ach = top["mb"] / top["ms"]     # MB/ms == GB/s
        roof = {"bound": "hbm", "achieved": round(ach, 1), "peak": round(BW / 1e9, 1), "unit": "GB/s",
                "frac": round(ach / (BW / 1e9), 3)}
    traffic = None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        for key, rec in tj.items():
            if not key.startswith("_") and top["op"].startswith(key) and N == CONFIGS[model][0]:
                traffic = {"bytes": rec["dram_read_bytes"] + rec["dram_write_bytes"], "algorithmic_bytes": top["mb"] * 1e6,
                           "source": rec["source"]}
    except (OSError, ValueError, KeyError):
        pass
    roof.update({"traffic": traffic, "kernel": top["op"], "kernel_ms": top["ms"], "peak_source": pk["source"],
                 "share_of_step": round(top["ms"] / sum_meas, 3)})
    tc = [r for r in table if r["op"].startswith("conv_tc") and r["bound"] == "tensor"]
    hb = [r for r in table if r["bound"] == "hbm"]
    res["roofline"] = roof
    res["roofline_step"] = {
        "sum_t_bound_ms": round(sum_bound, 3), "sum_measured_ms": round(sum_meas, 3),
        "frac": round(sum_bound / (ms_step if world == 1 else sum_meas), 3),
        "tensor_bound_convs": {"n": len(tc), "tflops": round(sum(r["gflop"] for r in tc) / max(sum(r["ms"] for r in tc), 1e-9), 1),
                               "min_frac": min((r["frac"] for r in tc), default=None)},
        "hbm_bound_ops": {"n": len(hb), "gbs": round(sum(r["mb"] for r in hb) / max(sum(r["ms"] for r in hb), 1e-9), 1)}}
    res["plan"] = {"ops": cm.num_ops, "launches_per_step": cm.num_launches, "arena_mb": round(cm.arena_bytes / 2 ** 20, 1),
                   "weights_mb": round(cm.weight_bytes / 2 ** 20, 1)}
    work_mb = x.numel() * 4 / 1e6 + cm.arena_bytes / 1e6
    res["l2"] = (("input batch (%.0f MB) + activation arena (%.0f MB) exceed the 126 MB L2; no explicit flush"
                  if work_mb > 126.0 else
                  "working set (%.1f MB input + %.1f MB arena) FITS the 126 MB L2 and is not flushed between steps: an "
                  "L2-warm steady-state number") % (x.numel() * 4 / 1e6, cm.arena_bytes / 1e6))
    try:
        out = a.ops_out if full else a.ops_out.replace(".json", f"_{model}.json")
        os.makedirs(os.path.dirname(out), exist_ok=True)
        json.dump({"model": model, "batch": N, "dtype": dtype, "ms_per_step": ms_step, "ops": table}, open(out, "w"), indent=1)
    except OSError:
        pass

    # ---------------- CPU baseline (rank 0, N = 1, headline only): the reference on the host cores, bounded sample ----------------
    if full and world == 1 and not a.no_cpu_baseline:
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        sample = min(N, a.ref_batch)
        xs = x_cpu[:sample]
        get_model, where = reference_get_model()
        if get_model is not None:
            rnet = build_net(model, H, W, get_model)
            kind = "reference"

            def fwd():
                with torch.no_grad():
                    return rnet(xs)
        else:
            from oracle import oracle_forward
            kind = "port"

            def fwd():
                return oracle_forward(net_cpu, xs)
        fwd()
        t0 = time.perf_counter()
        fwd()
        t1 = time.perf_counter() - t0
        reps_cpu = max(2, min(20, int(a.cpu_budget_s / max(t1, 1e-3))))
        t0 = time.perf_counter()
        for _ in range(reps_cpu):
            fwd()
        dt = (time.perf_counter() - t0) / reps_cpu
        res["cpu_baseline"] = {
            "value": round(sample / dt, 2), "unit": "images/s", "cores": cores, "kind": kind,
            "sample": f"{reps_cpu} timed forwards of {sample} of the {N} images after 2 warm-ups, fp32, torch "
                      f"{torch.__version__} CPU with {torch.get_num_threads()} threads"
                      + (" (unmodified pytorchcv module)" if kind == "reference" else " (oracle port)")}
    return res


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", type=str, default="b200", choices=["b200", "reference"])
    ap.add_argument("--model", type=str, default="resnet50", choices=sorted(CONFIGS))
    ap.add_argument("--batch", type=int, default=None, help="images per GPU (weak scaling) / global batch (strong scaling)")
    ap.add_argument("--dtype", type=str, default=None, choices=["bf16", "fp16", "fp32"])
    ap.add_argument("--scaling", type=str, default="weak", choices=["weak", "strong"])
    ap.add_argument("--exchange", type=str, default="auto", choices=["auto", "peer", "nccl"],
                    help="logits all-gather: device-initiated peer stores inside the plan, or NCCL")
    ap.add_argument("--graph", type=int, default=1, help="replay the plan from a CUDA graph")
    ap.add_argument("--ref-batch", type=int, default=64, help="images per forward of the cpu_baseline sample")
    ap.add_argument("--cpu-budget-s", type=float, default=12.0, help="CPU seconds spent on the cpu_baseline sample")
    ap.add_argument("--ref-budget-s", type=float, default=240.0, help="wall-clock bound of the --impl reference arm")
    ap.add_argument("--sustain-s", type=float, default=2.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the brief lines for the other BASELINE configs")
    ap.add_argument("--ops-out", type=str, default=os.path.join(ROOT, "gpurun_out", "bench_ops.json"))
    a = ap.parse_args()
    dflt_batch, a.h, a.w, dflt_dtype = CONFIGS[a.model]
    a.dtype = a.dtype or dflt_dtype
    world_env = int(os.environ.get("WORLD_SIZE", "1"))
    a.batch = a.batch or dflt_batch
    if a.scaling == "strong":
        if a.batch % max(world_env, 1) != 0:
            raise SystemExit(f"strong scaling needs batch {a.batch} divisible by {world_env} ranks")
        a.batch //= max(world_env, 1)
    a.warmup = max(a.warmup, 3)
    if a.impl == "reference":
        run_reference(a)
        return

    import torch
    import torch.distributed as dist
    from pytorchcv_b200 import parallel

    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a B200; the CUDA path has no CPU fallback")
    rank, world, local = parallel.init_from_env("nccl")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = bind_to_gpu_numa_node(local)
    pk = peaks()
    N, H, W, K, Wm = a.batch, a.h, a.w, a.steps, a.warmup

    r = measure(a, a.model, N, H, W, a.dtype, K, Wm, rank, world, local, dev, pk, full=True)

    others = []
    if world == 1 and not a.no_configs and a.model == "resnet50":
        for m in BASELINE_CONFIGS:
            if m == a.model:
                continue
            b, h, w, dt = CONFIGS[m]
            try:
                o = measure(a, m, b, h, w, dt, 50, Wm, rank, world, local, dev, pk, full=False)   # 50 timed steps each
                others.append({"config": workload_string(m, h, w, b), "model": m, "batch": b, "dtype": dt, "value": o["value"],
                               "unit": "images/s", "ms_per_step": o["ms_per_step"], "e2e": o["e2e"]["value"],
                               "e2e_f32": o["e2e_f32"]["value"], "roofline_step": o.get("roofline_step", {}).get("frac"),
                               "roofline": o.get("roofline"), "parity": o.get("parity"), "gpu_launches": o["gpu_launches"],
                               "clocks": o["clocks"]})
            except Exception as e:  # noqa: BLE001  (a secondary config must not take the headline line down)
                others.append({"model": m, "error": repr(e)[-300:]})
            torch.cuda.empty_cache()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    gb = N * world
    line = {
        "metric": f"{a.model} bs{N} {a.dtype} eval inference images/sec", "value": r["value"],
        "unit": "images/s", "n_gpus": world, "steps": K, "warmup": Wm, "ms_per_step": r["ms_per_step"],
        "higher_is_better": True, "scaling": a.scaling, "vs_baseline": None,
        "dtype": {"bf16": "bf16", "fp16": "f16", "fp32": "f32"}[a.dtype], "data": "synthetic",
        "config": {"workload": workload_string(a.model, H, W, N), "global_batch": gb,
                   "parallelism": (f"batch-sharded replicas x{world}, 1 all-gather of logits ({r['exchange']})" if world > 1
                                   else "single replica"),
                   "l2": r.get("l2"), "graph": bool(a.graph), "numa": numa},
        "clocks": r["clocks"], "e2e": r["e2e"], "e2e_f32": r["e2e_f32"], "gpu_launches": r["gpu_launches"],
        "parity": r.get("parity"), "sustained": r.get("sustained"),
        "roofline": r.get("roofline"), "roofline_step": r.get("roofline_step"), "cpu_baseline": r.get("cpu_baseline"),
        "plan": r.get("plan"),
    }
    if others:
        line["configs"] = others
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
