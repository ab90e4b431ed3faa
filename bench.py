#!/usr/bin/env python
"""bench.py — ResNet-50 bs256 bf16 eval-forward throughput on B200 (BASELINE.json's metric), one JSON line.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--model resnet50] [--batch 256] [--dtype bf16]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...     (N > 1)
    python bench.py --impl reference ...      # the reference's CPU path (oracle port) on the host cores

A "step" is one eval forward of the named network over one synthetic batch (weak scaling: `--batch` images per GPU).
`value` = images/s with inputs resident in HBM; `e2e` = images/s through the public module call with the batch in
pinned HOST memory (H2D of the fp32 NCHW batch and D2H of the logits inside the timed region, double-buffered).
`roofline` describes the single most expensive kernel launch of the step (per-op CUDA events); `roofline_step`
compares the whole step with the sum of per-kernel bounds.
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

CONFIGS = {  # model -> (default batch, H, W)
    "resnet50": (256, 224, 224), "resnet18": (8, 224, 224), "mobilenetv2_w1": (256, 224, 224),
    "seresnext50_32x4d": (256, 224, 224), "deeplabv3_resnetd50b_voc": (16, 480, 480),
    "efficientnet_b0": (256, 224, 224),   # SURVEY 8(f) rank 1 (not BASELINE configs: measured to the same bar)
    "mobilenetv3_large_w1": (256, 224, 224),
}


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
        self.index, self.lines, self.proc = index, [], None

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

    def summary(self) -> dict:
        sm, smax, pw, reasons = [], 0, 0.0, set()
        for ln in self.lines:
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


def build_net(model: str, H: int, W: int):
    import torch
    import pytorchcv_b200 as P
    torch.manual_seed(0)  # the reference's own random init (default BN), same on every rank
    kw = {"in_size": (H, W)} if model.startswith("deeplab") else {}
    return P.get_model(model, pretrained=False, **kw).eval()


def run_reference(a) -> None:
    """The reference's CPU implementation of the path (its torch-CPU forward, restated in oracle/), on host cores."""
    import torch
    from oracle import oracle_forward, seeded_input
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    batch, H, W = a.batch, a.h, a.w
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sample = min(batch, a.ref_batch)
    net = build_net(a.model, H, W)
    x = seeded_input((sample, 3, H, W), seed=1234)
    steps, warm = max(1, min(a.steps, a.ref_max_steps)), max(1, min(a.warmup, 2))
    for _ in range(warm):
        oracle_forward(net, x)
    t0 = time.perf_counter()
    for _ in range(steps):
        oracle_forward(net, x)
    dt = (time.perf_counter() - t0) / steps
    v = sample / dt
    # same metric / config strings as the b200 arm (the driver pairs the two lines); dtype says what this arm computes in
    line = {"impl": "reference", "metric": f"{a.model} bs{batch} {a.dtype} eval inference images/sec", "value": round(v, 2),
            "unit": "images/s", "n_gpus": a.gpus, "steps": steps, "warmup": warm, "ms_per_step": round(dt * 1e3, 2),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{a.model} eval forward, {H}x{W}, batch {batch} per GPU, fp32 NCHW in -> fp32 logits out "
                                   f"(random-init weights, torch.manual_seed(0))",
                       "global_batch": batch * a.gpus, "sample_batch": sample,
                       "parallelism": "reference CPU path on rank 0's host cores (torch fp32, all threads)"},
            "cpu_baseline": {"value": round(v, 2), "unit": "images/s", "cores": cores, "kind": "port",
                             "sample": f"{steps} timed forwards of {sample} images (torch {torch.__version__} CPU, "
                                       f"{torch.get_num_threads()} threads) through oracle/ref_forward.py"},
            "e2e": {"value": round(v, 2), "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", type=str, default="b200", choices=["b200", "reference"])
    ap.add_argument("--model", type=str, default="resnet50", choices=sorted(CONFIGS))
    ap.add_argument("--batch", type=int, default=None, help="images per GPU (weak scaling)")
    ap.add_argument("--dtype", type=str, default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--graph", type=int, default=1, help="replay the plan from a CUDA graph")
    ap.add_argument("--ref-batch", type=int, default=32, help="images per step of the CPU reference arm")
    ap.add_argument("--ref-max-steps", type=int, default=6)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ops-out", type=str, default=os.path.join(ROOT, "gpurun_out", "bench_ops.json"))
    a = ap.parse_args()
    dflt_batch, a.h, a.w = CONFIGS[a.model]
    a.batch = a.batch or dflt_batch
    a.warmup = max(a.warmup, 3)
    if a.impl == "reference":
        run_reference(a)
        return

    import torch
    import torch.distributed as dist
    import pytorchcv_b200 as P
    from pytorchcv_b200 import _lib, parallel

    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a B200; the CUDA path has no CPU fallback")
    rank, world, local = parallel.init_from_env("nccl")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    N, H, W, K, Wm = a.batch, a.h, a.w, a.steps, a.warmup
    pk = peaks()

    net = build_net(a.model, H, W).to(dev)
    fast = P.accelerate(net, dtype=a.dtype, graph=bool(a.graph), check_weights=False)
    x = torch.randn(N, 3, H, W, generator=torch.Generator().manual_seed(1234 + rank)).to(dev)
    runner = parallel.ShardedInference(fast, rank, world)

    def first(y):
        return y[0] if isinstance(y, (tuple, list)) else y

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local])
        torch.cuda.synchronize(dev)

    # ---------------- value: inputs resident in HBM ----------------
    y = runner(x)
    cm = fast.compiled(x)
    for _ in range(Wm):
        runner(x)
    barrier()
    l0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:
        barrier()
        e0.record()
        for _ in range(K):
            y = runner(x)
        e1.record()
        barrier()
    launches = _lib.launch_count() - l0
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_step = ms.item() / K
    value = N * world / (ms_step * 1e-3)

    # ---------------- e2e: host batch -> H2D -> forward -> D2H logits, double-buffered ----------------
    out0 = first(y)
    xh = [torch.randn(N, 3, H, W, generator=torch.Generator().manual_seed(99 + rank + i)).pin_memory() for i in range(2)]
    xd = [torch.empty_like(x) for _ in range(2)]
    out_shape = tuple(first(runner(x)).shape)  # [N*world, classes] once the logits are gathered
    yh = [torch.empty(out_shape, dtype=torch.float32).pin_memory() for _ in range(2)]
    comp, copy = torch.cuda.current_stream(dev), torch.cuda.Stream(device=dev)
    copied = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]

    def e2e_step(i):
        b = i & 1
        with torch.cuda.stream(copy):
            copy.wait_event(consumed[b])
            xd[b].copy_(xh[b], non_blocking=True)
            copied[b].record(copy)
        comp.wait_event(copied[b])
        out = first(runner(xd[b]))
        consumed[b].record(comp)
        yh[b].copy_(out, non_blocking=True)

    for i in range(Wm):
        e2e_step(i)
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for i in range(K):
        e2e_step(i)
    f1.record()
    barrier()
    ems = torch.tensor([f0.elapsed_time(f1)], device=dev)
    if world > 1:
        dist.all_reduce(ems, op=dist.ReduceOp.MAX)
    e2e_value = N * world / (ems.item() / K * 1e-3)
    h2d = x.numel() * 4
    d2h = yh[0].numel() * 4

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

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
            if not key.startswith("_") and top["op"].startswith(key) and N == CONFIGS[a.model][0]:
                traffic = {"bytes": rec["dram_read_bytes"] + rec["dram_write_bytes"], "algorithmic_bytes": top["mb"] * 1e6,
                           "source": rec["source"]}
    except (OSError, ValueError, KeyError):
        pass
    roof.update({"traffic": traffic, "kernel": top["op"], "kernel_ms": top["ms"], "peak_source": pk["source"],
                 "share_of_step": round(top["ms"] / sum_meas, 3)})
    tc = [r for r in table if r["op"].startswith("conv_tc") and r["bound"] == "tensor"]
    hb = [r for r in table if r["bound"] == "hbm"]
    roof_step = {"sum_t_bound_ms": round(sum_bound, 3), "sum_measured_ms": round(sum_meas, 3),
                 "frac": round(sum_bound / (ms_step if world == 1 else sum_meas), 3),
                 "tensor_bound_convs": {"n": len(tc), "tflops": round(sum(r["gflop"] for r in tc) / max(sum(r["ms"] for r in tc), 1e-9), 1)},
                 "hbm_bound_ops": {"n": len(hb), "gbs": round(sum(r["mb"] for r in hb) / max(sum(r["ms"] for r in hb), 1e-9), 1)}}
    try:
        os.makedirs(os.path.dirname(a.ops_out), exist_ok=True)
        json.dump({"model": a.model, "batch": N, "dtype": a.dtype, "ms_per_step": ms_step, "ops": table},
                  open(a.ops_out, "w"), indent=1)
    except OSError:
        pass

    # ---------------- CPU baseline (rank 0, N = 1 only): the oracle port on the host cores ----------------
    cpu = None
    if world == 1 and not a.no_cpu_baseline:
        from oracle import oracle_forward, seeded_input
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        sample = min(N, a.ref_batch)
        xs = seeded_input((sample, 3, H, W), seed=1234)
        cnet = build_net(a.model, H, W)
        oracle_forward(cnet, xs)
        t0 = time.perf_counter()
        reps_cpu = 2
        for _ in range(reps_cpu):
            oracle_forward(cnet, xs)
        dt = (time.perf_counter() - t0) / reps_cpu
        cpu = {"value": round(sample / dt, 2), "unit": "images/s", "cores": cores, "kind": "port",
               "sample": f"{reps_cpu} timed forwards of {sample} of the {N} images after 1 warm-up, fp32, "
                         f"torch {torch.__version__} CPU with {torch.get_num_threads()} threads"}

    work_mb = h2d / 1e6 + cm.arena_bytes / 1e6
    if work_mb > 126.0:
        l2_note = ("input batch (%.0f MB) + activation arena (%.0f MB) exceed the 126 MB L2; no explicit flush"
                   % (h2d / 1e6, cm.arena_bytes / 1e6))
    else:
        l2_note = ("working set (%.1f MB input + %.1f MB arena) FITS the 126 MB L2 and is not flushed between steps: "
                   "an L2-warm steady-state number" % (h2d / 1e6, cm.arena_bytes / 1e6))
    line = {
        "metric": f"{a.model} bs{N} {a.dtype} eval inference images/sec", "value": round(value, 1),
        "unit": "images/s", "n_gpus": world, "steps": K, "warmup": Wm, "ms_per_step": round(ms_step, 4),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": a.dtype if a.dtype == "bf16" else "f32", "data": "synthetic",
        "config": {"workload": f"{a.model} eval forward, {H}x{W}, batch {N} per GPU, fp32 NCHW in -> fp32 logits out "
                               f"(random-init weights, torch.manual_seed(0))",
                   "global_batch": N * world, "parallelism": f"batch-sharded replicas x{world}, 1 all-gather of logits",
                   "l2": l2_note,
                   "graph": bool(a.graph)},
        "clocks": clk.summary(),
        "e2e": {"value": round(e2e_value, 1), "unit": "images/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": round(ems.item() / K, 4), "pipelining": "2 pinned host + 2 device buffers, copy stream"},
        "gpu_launches": int(launches),
        "roofline": roof, "roofline_step": roof_step, "cpu_baseline": cpu,
        "plan": {"ops": cm.num_ops, "launches_per_step": cm.num_launches, "arena_mb": round(cm.arena_bytes / 2 ** 20, 1),
                 "weights_mb": round(cm.weight_bytes / 2 ** 20, 1)},
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
