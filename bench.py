#!/usr/bin/env python
"""Benchmark of the hot path: 1024x1024 tiles/s, forward + backward (+ fused AdamW step) of the ResNeSt-50 U-Net
(BASELINE.json config 2: --type pre, batch 8 per GPU, bf16, focal+dice) on N B200s, one process per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W]                 # ours (libxv2, hand-written sm_100a kernels)
    python bench.py --impl reference [--gpus N] [--steps K] [--warmup W]  # the reference's CPU PyTorch path (oracle port)
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...    # N > 1 (the driver launches it this way)

Prints ONE JSON line on rank 0.  Keys beyond the base contract: `roofline` (dominant kernel, measured with CUDA events
in an instrumented step after the timed region), `cpu_baseline`, `e2e` (pinned host uint8 tiles -> H2D on a side stream
-> step -> loss read back, all inside the timed region), `clocks`, `gpu_launches`, `kernels` (per entry-point split).
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

METRIC = "1024x1024 tiles/sec (fwd+bwd) ResNeSt50-UNet"
UNIT = "tiles/s"
FWD_BWD_GFLOP_PER_TILE = 1856.7  # conv/convT/GEMM FLOPs, ResNeSt-50 U-Net, one 1024^2 tile, fwd+bwd


def config_namespace(a):
    return argparse.Namespace(
        ppm=False, aspp=False, dilation=1, no_skip=False, interpolate=False, attention=False, dec_interp=False,
        deep_supervision=False, loss_str="focal+dice", encoder=a.encoder, dmg_model="siamese", type="pre", tta=False,
        precision="bf16", lr=3e-4, optimizer="adamw", weight_decay=0.0, momentum=0.9, use_scheduler=False, warmup=1,
        epochs=1, gpus=a.gpus, init_lr=1e-4, final_lr=1e-4, results=None, logname="bench", autoaugment=False)


def workload_config(batch, size, world):
    """The `config` object of BOTH arms (ours and --impl reference): BASELINE.json configs[1]."""
    return {"workload": f"ResNeSt-50 U-Net --type pre, batch {batch}/GPU, {size}x{size}x3 synthetic tiles, bf16 compute / fp32 master "
                        f"weights, focal+dice loss, forward + backward + fused AdamW step (BASELINE.json configs[1])",
            "global_batch": batch * world,
            "parallelism": f"dp{world}: tiles sharded over ranks, one NCCL all-reduce of the flat gradient buffer",
            "l2": "activations per step (tens of GB) exceed the 126 MB L2; no explicit flush needed",
            "launch": "forward + backward replayed from one CUDA graph; gradient all-reduce, fused AdamW and weight re-pack eager"}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return {"tensor": d["bf16_tflops_sustained"], "tensor_burst": d["bf16_tflops"], "hbm": d["hbm_gbs"], "src": "measured"}
    return {"tensor": 1400.0, "tensor_burst": 1590.0, "hbm": 6650.0, "src": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) < 7:
                continue
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                power.append(float(r[2]))
            except ValueError:
                continue
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------------------------------------
# the reference's CPU path (oracle port): used by `cpu_baseline` and by --impl reference
# ---------------------------------------------------------------------------------------------------------------
CPU_SAMPLE = {"batch": 2, "size": 512}  # 2 x 512^2 = half a 1024^2 tile of pixels per step


def cpu_reference_step_fn(encoder):
    import torch

    from oracle import functional as OF
    torch.set_num_threads(os.cpu_count() or 1)
    ns = argparse.Namespace(encoder=encoder, attention=False, deep_supervision=False, type="pre", dmg_model="siamese",
                            loss_str="focal+dice")
    # parameter shapes from the product model's state_dict (construction only, no compute), values from the seeded fill
    from xview2_b200.model.unet import UNetLoc
    shapes = {k: (tuple(v.shape), v.dtype) for k, v in UNetLoc(config_namespace(argparse.Namespace(encoder=encoder, gpus=1))).state_dict().items()}
    state = OF.deterministic_state(shapes, 1)
    g = torch.Generator().manual_seed(1)
    b, s = CPU_SAMPLE["batch"], CPU_SAMPLE["size"]
    x = torch.randn(b, 3, s, s, generator=g)
    y = torch.randint(0, 2, (b, s, s), generator=g, dtype=torch.uint8)
    P = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running_" not in k else v.clone()) for k, v in state.items()}

    def step():
        for v in P.values():
            if v.grad is not None:
                v.grad = None
        out = OF.model_forward(P, x, True, ns)
        loss = OF.compute_loss(out, y, ns.loss_str, False, False)
        loss.backward()
        return float(loss.detach())

    tiles_per_step = b * s * s / (1024.0 * 1024.0)
    return step, tiles_per_step


def time_cpu_reference(encoder, steps, warmup):
    step, tiles = cpu_reference_step_fn(encoder)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / max(1, steps)
    return tiles / dt, dt


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    value, dt = time_cpu_reference(a.encoder, a.steps, a.warmup)
    cores = os.cpu_count() or 1
    sample = (f"oracle port (plain PyTorch fp32, {cores} threads): ResNeSt-50 U-Net fwd+bwd on {CPU_SAMPLE['batch']} x "
              f"{CPU_SAMPLE['size']}^2 synthetic crops per step = {CPU_SAMPLE['batch'] * CPU_SAMPLE['size'] ** 2 / 1024 ** 2:.2f} "
              f"tile-equivalents of pixels")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": dict({k: v for k, v in workload_config(a.batch, a.size, max(1, a.gpus)).items() if k != "launch"},
                           reference_impl="CPU PyTorch path of the reference, fp32, all host cores (oracle port: /root/reference is a "
                                          "script tree whose pytorch_lightning / apex / monai / resnest imports cannot be installed)",
                           reference_sample=sample),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------------
# ours
# ---------------------------------------------------------------------------------------------------------------
def run_ours(a):
    import torch
    import torch.distributed as dist

    from xview2_b200 import lib, ops
    from xview2_b200.data_loading.ring import TileRing
    from xview2_b200.model.plt import Model

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # the contract is ONE JSON line on stdout: native libraries (NCCL prints its version banner on fd 1) go to stderr
    sys.stdout.flush()
    json_out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py measures the CUDA path; no GPU is visible (there is no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        import datetime
        dist.init_process_group("nccl", device_id=torch.device("cuda", local), timeout=datetime.timedelta(seconds=180))
    lib.init(local)
    ops.enable_wgrad_side_stream(True)
    dev = torch.device("cuda", local)
    B, S = a.batch, a.size

    ns = config_namespace(a)
    torch.manual_seed(1)
    model = Model(ns).to(dev).train()
    opt = model.configure_optimizers()
    flat = model.flat
    flat.broadcast_params(0)

    # synthetic tiles: seeded uint8 "decoded PNG" bytes in the pinned ring; blocky labels
    ring = TileRing(B, S, S, post=False, slots=2, device=dev)
    g = torch.Generator().manual_seed(1 + rank)
    for i in range(ring.slots):
        slot = ring.host(i)
        slot["tiles"].copy_(torch.randint(0, 256, slot["tiles"].shape, generator=g, dtype=torch.uint8))
        cells = torch.randint(0, 2, (B, S // 32, S // 32), generator=g, dtype=torch.uint8)
        slot["mask"].copy_(cells.repeat_interleave(32, 1).repeat_interleave(32, 2))
    resident = {k: v.to(dev) for k, v in ring.host(0).items()}

    def train_step(batch):
        opt.zero_grad()
        loss = model.training_step(batch, 0)
        loss.backward()
        n = flat.all_reduce_grads()
        opt.grad_scale = 1.0 / n
        opt.step()
        return loss

    def fence():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    eager_step = train_step
    use_graph = os.environ.get("XV2_NO_GRAPH", "0") != "1"
    if use_graph:
        # forward + backward captured once in a CUDA graph and replayed; all-reduce + fused AdamW + re-pack stay eager
        from xview2_b200.graph import GraphedTrainStep
        try:
            gstep = GraphedTrainStep(model, opt, resident, warmup=2)
            train_step = gstep  # same signature: batch -> loss tensor
        except Exception as exc:  # noqa: BLE001 -- the capture is an optimisation: the eager step measures the same work
            print(f"[bench] CUDA-graph capture failed ({type(exc).__name__}: {exc}); timing the eager step", file=sys.stderr)
            torch.cuda.synchronize()
            use_graph = False

    # ---- device-resident throughput ("value") ------------------------------------------------------------------
    for _ in range(a.warmup):
        train_step(resident)
    fence()
    sampler = ClockSampler(local) if rank == 0 else None
    launches0 = lib.launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        loss = train_step(resident)
    e1.record()
    fence()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms)
    launches = lib.launches() - launches0
    last_loss = float(loss.detach())

    # ---- end to end: pinned host tiles -> side-stream H2D -> step -> loss read back, every step ---------------------
    ring.submit(0)
    for i in range(2):  # warm the ring path
        ring.submit(i + 1)
        batch = ring.acquire(i)
        float(train_step(batch).detach())
        ring.release(i)
    fence()
    # The loss of EVERY step is read back on the host inside the timed region, one step behind the launch front (the read of
    # step i is issued after step i+1 has been enqueued), so the device never waits for Python between steps.
    pinned_loss = torch.empty(a.steps, dtype=torch.float32).pin_memory()
    host_losses = []
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    pending = None
    for j, i in enumerate(range(2, 2 + a.steps)):
        ring.submit(i + 1)           # next batch's copy overlaps this step
        batch = ring.acquire(i)
        l = train_step(batch)
        ring.release(i)
        pinned_loss[j:j + 1].copy_(l.detach().reshape(1), non_blocking=True)  # D2H of this step's result
        done = torch.cuda.Event()
        done.record()
        if pending is not None:
            pending[1].synchronize()
            host_losses.append(float(pinned_loss[pending[0]]))
        pending = (j, done)
    pending[1].synchronize()
    host_losses.append(float(pinned_loss[pending[0]]))
    e3.record()
    fence()
    clocks = sampler.stop() if sampler else None
    ms2 = torch.tensor([e2.elapsed_time(e3)], device=dev)
    if world > 1:
        dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
    e2e_value = B * world * a.steps / (float(ms2) / 1e3)

    # ---- roofline of the dominant kernel: instrumented step (CUDA events around every libxv2 launch) -------------
    # EVERY rank runs the step (it contains the gradient all-reduce); only rank 0 records the events
    roof = kernels = None
    if rank == 0:
        lib.profile_start()
    eager_step(resident)
    fence()
    if rank == 0:
        prof = lib.profile_stop()
        total = sum(d["ms"] for d in prof.values()) or 1.0
        kernels = {k: {"calls": d["calls"], "ms": round(d["ms"], 3), "share": round(d["ms"] / total, 4),
                       "tflops": round(d["flops"] / d["ms"] / 1e9, 1) if d["flops"] and d["ms"] else None}
                   for k, d in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])}
        top = max(prof.items(), key=lambda kv: kv[1]["ms"])
        peaks = measured_peaks()
        name, d = top
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):  # ncu dram__bytes_read + write per launch of this entry point's kernels (tools/make_traffic.py)
            with open(tpath) as f:
                traffic = (json.load(f).get(name) or {}).get("dram_bytes_per_launch")
        if d["flops"] > 0:
            ach = d["flops"] / d["ms"] / 1e9  # TFLOP/s over all launches of this entry point in the step
            roof = {"kernel": name, "bound": "tensor", "achieved": round(ach, 1), "peak": peaks["tensor"], "unit": "TFLOP/s",
                    "frac": round(ach / peaks["tensor"], 4), "traffic": traffic, "peak_source": peaks["src"] + " (sustained bf16)",
                    "launches": d["calls"], "avg_launch_ms": round(d["ms"] / d["calls"], 4), "share_of_step": round(d["ms"] / total, 4),
                    "algorithmic_gflop_per_launch": round(d["flops"] / d["calls"] / 1e9, 2),
                    "algorithmic_bytes_per_launch": round(d["bytes"] / d["calls"]),
                    "hbm_gbs": round(d["bytes"] / d["ms"] / 1e6, 1), "hbm_frac": round(d["bytes"] / d["ms"] / 1e6 / peaks["hbm"], 4),
                    "note": "the conv entry point runs on the ridge: both the tensor and the HBM fraction are reported"}
        else:
            ach = d["bytes"] / d["ms"] / 1e6  # GB/s
            roof = {"kernel": name, "bound": "hbm", "achieved": round(ach, 1), "peak": peaks["hbm"], "unit": "GB/s",
                    "frac": round(ach / peaks["hbm"], 4), "traffic": traffic, "peak_source": peaks["src"],
                    "launches": d["calls"], "avg_launch_ms": round(d["ms"] / d["calls"], 4), "share_of_step": round(d["ms"] / total, 4)}

    cpu_baseline = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        v, dt = time_cpu_reference(a.encoder, 2, 1)
        cores = os.cpu_count() or 1
        cpu_baseline = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                        "sample": f"oracle port fwd+bwd on {CPU_SAMPLE['batch']} x {CPU_SAMPLE['size']}^2 crops "
                                  f"(0.5 tile-equivalents), 1 warm-up + 2 timed steps, {dt:.1f} s/step"}

    if rank == 0:
        value = B * world * a.steps / (ms_total / 1e3)
        peaks = measured_peaks()
        cfg = workload_config(B, S, world)
        if not use_graph:
            cfg["launch"] = "eager launches (no CUDA graph)"
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms_total / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": cfg,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": ring.bytes_per_batch, "d2h_bytes_per_step": 4,
                    "path": "pinned uint8 HWC tiles + masks -> side-stream H2D (double-buffered TileRing) -> Model.training_step -> backward -> all-reduce -> AdamW -> loss D2H into pinned memory, read on the host every step (one step behind the launch front)",
                    "losses_read": len(host_losses)},
            "gpu_launches": launches,
            "conv_roofline_frac": round(FWD_BWD_GFLOP_PER_TILE * value / 1e3 / peaks["tensor"], 4),
            "roofline": roof, "cpu_baseline": cpu_baseline, "clocks": clocks, "kernels": kernels, "loss": last_loss,
        }
        json_out.write(json.dumps(line) + "\n")
        json_out.flush()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=8, help="tiles per GPU per step (BASELINE configs[1]: 8)")
    ap.add_argument("--size", type=int, default=1024)
    ap.add_argument("--encoder", default="resnest50")
    ap.add_argument("--no-cpu-baseline", dest="no_cpu_baseline", action="store_true")
    a = ap.parse_args()
    if a.warmup < 3 and a.impl == "ours":
        a.warmup = 3
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)


if __name__ == "__main__":
    main()
