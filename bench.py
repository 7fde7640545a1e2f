#!/usr/bin/env python
"""Benchmark of the hot path on N B200s, one process per GPU.

    python bench.py [--config c2|c3|c4|c5] [--gpus N] [--steps K] [--warmup W]    # ours (libxv2, hand-written sm_100a kernels)
    python bench.py --impl reference ...   # the reference's CPU PyTorch path (oracle port) on the host cores
    python bench.py --impl library ...     # the reference's modules (oracle port) on the SAME B200 under stock PyTorch /
                                           # cuDNN: bf16 autocast, channels_last, cudnn.benchmark (main.py:96-111) -- "the bar"
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...              # N > 1 (the driver launches it this way)

Configurations (BASELINE.json `configs`; the headline metric is quoted on c2 = configs[1], the default):
    c2  ResNeSt-50 U-Net --type pre, batch 8/GPU, focal+dice, forward + backward + fused AdamW           tiles/s
    c3  ResNeSt-101 Siamese U-Net --type post, batch 4 pairs/GPU                                         pairs/s
    c4  ResNeSt-200 fused U-Net + deep supervision + attention, batch 2 pairs/GPU                        pairs/s
    c5  ResNeSt-50 U-Net eval: 4-pass TTA + argmax label map + F1 counters, batch 16/GPU                 tiles/s

Prints ONE JSON line on rank 0.  Keys beyond the base contract: `roofline` (dominant entry point, CUDA events in an
instrumented step after the timed region; `roofline.stages` = per-stage forward fractions), `cpu_baseline` (+ `cpu_baseline_c1`:
SURVEY 8d's C1 exactly), `library_baseline`, `e2e` (pinned host uint8 tiles -> H2D on a side stream -> step -> result read
back, all inside the timed region), `clocks`, `gpu_launches`, `kernels` (per entry-point split).
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

# SURVEY.md 8(d) algorithmic GFLOP per unit (conv / convT / GEMM only; backward = dgrad + wgrad)
CONFIGS = {
    "c2": dict(encoder="resnest50", type="pre", dmg_model="siamese", batch=8, deep_supervision=False, attention=False,
               tta=False, mode="train", unit="tiles/s", gflop=1856.7, metric=METRIC, baseline="BASELINE.json configs[1]",
               what="ResNeSt-50 U-Net --type pre"),
    "c3": dict(encoder="resnest101", type="post", dmg_model="siamese", batch=4, deep_supervision=False, attention=False,
               tta=False, mode="train", unit="pairs/s", gflop=5204.3, baseline="BASELINE.json configs[2]",
               metric="1024x1024 pre/post pairs/sec (fwd+bwd) ResNeSt101 Siamese U-Net",
               what="ResNeSt-101 Siamese U-Net --type post --dmg_model siamese"),
    "c4": dict(encoder="resnest200", type="post", dmg_model="fused", batch=2, deep_supervision=True, attention=True,
               tta=False, mode="train", unit="pairs/s", gflop=13207.3, baseline="BASELINE.json configs[3]",
               metric="1024x1024 pre/post pairs/sec (fwd+bwd) ResNeSt200 fused U-Net + deep supervision + attention",
               what="ResNeSt-200 fused U-Net --type post --dmg_model fused --deep_supervision --attention"),
    "c5": dict(encoder="resnest50", type="pre", dmg_model="siamese", batch=16, deep_supervision=False, attention=False,
               tta=True, mode="eval", unit="tiles/s", gflop=2315.2, baseline="BASELINE.json configs[4]",
               metric="1024x1024 tiles/sec (eval, 4-pass TTA + argmax) ResNeSt50-UNet",
               what="ResNeSt-50 U-Net --type pre --exec_mode eval --tta + argmax label map + F1 counters"),
}


def resolve_config(a):
    c = dict(CONFIGS[a.config])
    if a.batch is not None:
        c["batch"] = a.batch
    if a.encoder is not None:
        c["encoder"] = a.encoder
    c["size"] = a.size
    c["post"] = c["type"] == "post"
    return c


def config_namespace(a, c=None):
    """argparse.Namespace a reference user would have after `main.py` parsed the flags of this configuration."""
    c = c or dict(CONFIGS["c2"], encoder=getattr(a, "encoder", None) or "resnest50")
    return argparse.Namespace(
        ppm=False, aspp=False, dilation=1, no_skip=False, interpolate=False, attention=c["attention"], dec_interp=False,
        deep_supervision=c["deep_supervision"], loss_str="focal+dice", encoder=c["encoder"], dmg_model=c["dmg_model"],
        type=c["type"], tta=c["tta"], precision="bf16", lr=3e-4, optimizer="adamw", weight_decay=0.0, momentum=0.9,
        use_scheduler=False, warmup=1, epochs=1, gpus=getattr(a, "gpus", 1), init_lr=1e-4, final_lr=1e-4, results=None,
        logname="bench", autoaugment=False)


def workload_config(c, world):
    """The `config` object -- IDENTICAL in every arm (ours, --impl reference, --impl library); what each arm actually ran
    per step is stated in its own `sample` / `launch` keys outside `config`."""
    unit = "pre/post pairs" if c["post"] else "tiles"
    step = ("forward + backward + AdamW step" if c["mode"] == "train"
            else "eval forward x4 (TTA flips) + logit mean + argmax label map + F1 counters")
    return {"workload": f"{c['what']}, batch {c['batch']}/GPU, {c['size']}x{c['size']}x3 synthetic {unit}, focal+dice loss, "
                        f"{step} ({c['baseline']})",
            "global_batch": c["batch"] * world,
            "parallelism": f"dp{world}: {unit} sharded over ranks, NCCL all-reduce of the flat gradient buffer only "
                           f"(per-stage buckets overlapped with backward)",
            "l2": "activations per step (tens of GB) exceed the 126 MB L2; no explicit flush needed"}


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
# the reference's modules (oracle port): CPU arm (`cpu_baseline`, --impl reference) and library arm (--impl library)
# ---------------------------------------------------------------------------------------------------------------
def oracle_state(c, device="cpu"):
    """Seeded parameters for the oracle's functional model: shapes from the product model's state_dict (construction only,
    no compute), values from the oracle's deterministic fill."""
    import torch

    from oracle import functional as OF
    from xview2_b200.model.plt import Model
    ns = config_namespace(argparse.Namespace(gpus=1), c)
    shapes = {k[len("model."):]: (tuple(v.shape), v.dtype) for k, v in Model(ns).state_dict().items() if k.startswith("model.")}
    state = OF.deterministic_state(shapes, 1)
    P = {}
    for k, v in state.items():
        if OF.canonical_key(k) != k:  # alias spelling of a shared module (FusedUNet): the oracle reads the canonical key
            continue
        v = v.to(device)
        if v.dim() == 4 and device != "cpu":
            v = v.contiguous(memory_format=torch.channels_last)
        if v.is_floating_point() and "running_" not in k:
            v.requires_grad_(c["mode"] == "train")
        P[k] = v
    return P, ns


def synthetic_float_batch(c, batch, size, seed=1):
    """What the reference's loader yields (pytorch_loader.py:163-171): float32 normalised CHW images + uint8 masks."""
    import torch
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(batch, 6 if c["post"] else 3, size, size, generator=g)
    cells = torch.randint(0, 5 if c["post"] else 2, (batch, max(1, size // 32), max(1, size // 32)), generator=g, dtype=torch.uint8)
    y = cells.repeat_interleave(32, 1).repeat_interleave(32, 2)[:, :size, :size].contiguous()
    return x, y


def oracle_step_fn(c, P, ns, x, y, optimizer=None, autocast=False):
    import torch

    from oracle import functional as OF
    dev_type = x.device.type

    def train():
        if optimizer is not None:
            optimizer.zero_grad(set_to_none=True)
        else:
            for v in P.values():
                v.grad = None
        with torch.autocast(dev_type, dtype=torch.bfloat16, enabled=autocast):
            out = OF.model_forward(P, x, True, ns)
        out = [o.float() for o in out] if isinstance(out, (list, tuple)) else out.float()
        loss = OF.compute_loss(out, y, ns.loss_str, c["post"], ns.deep_supervision)
        loss.backward()
        if optimizer is not None:
            optimizer.step()
        return loss.detach()

    def evaluate():
        with torch.no_grad(), torch.autocast(dev_type, dtype=torch.bfloat16, enabled=autocast):
            pred = OF.tta_forward(P, x, ns) if ns.tta else OF.model_forward(P, x, False, ns)
            lab = pred.float().argmax(1)
            t = y.long()
            tp, fp, fn = ((lab == 1) & (t == 1)).sum(), ((lab == 1) & (t != 1)).sum(), ((lab != 1) & (t == 1)).sum()
        return torch.stack([tp, fp, fn]).float().sum()

    return train if c["mode"] == "train" else evaluate


CPU_BUDGET_NOTE = "bounded sample so that --steps K --warmup W ends within minutes on the host cores"


def cpu_sample(c):
    """(batch, size) of one CPU step: c2 runs UNSCALED 1024^2 tiles at the smallest batch train-mode ResNeSt accepts (2: the
    split-attention BatchNorm needs > 1 sample); the larger configurations run 2 x 512^2 crops and are scaled by pixel count."""
    return (2, c["size"]) if (c["encoder"] == "resnest50" and c["mode"] == "train") else (2, min(512, c["size"]))


def time_cpu_reference(c, steps, warmup):
    import torch
    torch.set_num_threads(os.cpu_count() or 1)
    b, s = cpu_sample(c)
    P, ns = oracle_state(c)
    x, y = synthetic_float_batch(c, b, s)
    step = oracle_step_fn(c, P, ns, x, y)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / max(1, steps)
    units = b * s * s / float(c["size"] * c["size"])
    what = "pre/post pairs" if c["post"] else "tiles"
    sample = (f"oracle port (reference modules restated 1:1 in plain PyTorch, fp32, {torch.get_num_threads()} threads): "
              f"{c['what']} {'fwd+bwd' if c['mode'] == 'train' else 'eval TTA'} on {b} x {s}^2 synthetic {what} per step = "
              f"{units:.2f} full-size {what}; {dt:.2f} s/step; {CPU_BUDGET_NOTE}")
    return units / dt, dt, sample


def time_cpu_c1(steps=3, warmup=1):
    """SURVEY.md 8(d) / BASELINE.md section 4, C1 exactly: ResNet-50 U-Net --type pre, bs 1, one synthetic 1024^2 tile (seed 1),
    fp32, model.train(), forward + loss.backward(); 1 warm-up + >= 3 timed iterations, median."""
    import torch
    torch.set_num_threads(os.cpu_count() or 1)
    c = dict(CONFIGS["c2"], encoder="resnet50", batch=1, size=1024, post=False)
    P, ns = oracle_state(c)
    x, y = synthetic_float_batch(c, 1, 1024)
    step = oracle_step_fn(c, P, ns, x, y)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        step()
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    times.sort()
    med = times[len(times) // 2]
    return {"value": 1.0 / med, "unit": "tiles/s", "cores": os.cpu_count() or 1, "threads": torch.get_num_threads(), "kind": "port",
            "s_per_tile_median": round(med, 3),
            "sample": f"C1 exactly: ResNet-50 U-Net --type pre, bs 1, one randn(1,3,1024,1024) tile, fp32, train mode, fwd + "
                      f"loss.backward(); {warmup} warm-up + {steps} timed, median",
            "parallel_info": torch.__config__.parallel_info().split("\n")[0:3]}


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    c = resolve_config(a)
    value, dt, sample = time_cpu_reference(c, a.steps, a.warmup)
    cores = os.cpu_count() or 1
    line = {"impl": "reference", "metric": c["metric"], "value": value, "unit": c["unit"], "n_gpus": a.gpus, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": workload_config(c, max(1, a.gpus)),
            "reference_impl": "CPU PyTorch path of the reference, fp32, all host cores (oracle port: /root/reference is a script "
                              "tree whose pytorch_lightning / apex / monai / resnest imports cannot be installed offline)",
            "cpu_baseline": {"value": value, "unit": c["unit"], "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": c["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def time_library(c, steps, warmup, device, e2e=True):
    """The bar on this box: the reference's modules (oracle port) under stock PyTorch / cuDNN with the settings main.py:96-111
    selects (AMP -> bf16 autocast here, cudnn.benchmark=True), channels_last, fused torch AdamW.  Returns a dict."""
    import torch
    torch.backends.cudnn.benchmark = True
    b, s = c["batch"], c["size"]
    P, ns = oracle_state(c, device)
    xh, yh = synthetic_float_batch(c, b, s)
    xh, yh = xh.contiguous(memory_format=torch.channels_last).pin_memory(), yh.pin_memory()
    x, y = xh.to(device), yh.to(device)
    params = [v for v in P.values() if v.requires_grad]
    opt = torch.optim.AdamW(params, lr=3e-4, weight_decay=0.0, fused=True) if c["mode"] == "train" else None
    step = oracle_step_fn(c, P, ns, x, y, opt, autocast=True)
    for _ in range(max(3, warmup)):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    out = {"value": b / ms * 1e3, "unit": c["unit"], "ms_per_step": ms, "kind": "library",
           "peak_mem_gib": round(torch.cuda.max_memory_allocated() / 2 ** 30, 1),
           "sample": f"oracle port of the reference modules on this GPU: torch {torch.__version__} / cuDNN {torch.backends.cudnn.version()}, "
                     f"bf16 autocast, channels_last, cudnn.benchmark=True, "
                     f"{'fused torch.optim.AdamW, ' if opt is not None else ''}batch {b} x {s}^2, full {c['what']}"}
    if e2e:  # the reference's loader hands float32 CHW batches to .to(device) (pytorch_loader.py:163-171): 12 B / pixel / image
        xs = torch.empty_like(x)
        ys = torch.empty_like(y)
        step2 = oracle_step_fn(c, P, ns, xs, ys, opt, autocast=True)
        torch.cuda.synchronize()
        e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e2.record()
        for _ in range(steps):
            xs.copy_(xh, non_blocking=True)
            ys.copy_(yh, non_blocking=True)
            float(step2())
        e3.record()
        torch.cuda.synchronize()
        out["e2e"] = {"value": b * steps / (e2.elapsed_time(e3) / 1e3), "unit": c["unit"],
                      "h2d_bytes_per_step": xh.numel() * 4 + yh.numel(), "d2h_bytes_per_step": 4}
    del P, opt, params
    torch.cuda.empty_cache()
    return out


def run_library(a):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    c = resolve_config(a)
    torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    sampler = ClockSampler(0)
    r = time_library(c, a.steps, a.warmup, dev)
    clocks = sampler.stop()
    peaks = measured_peaks()
    line = {"impl": "library", "metric": c["metric"], "value": r["value"], "unit": c["unit"], "n_gpus": 1, "steps": a.steps,
            "warmup": max(3, a.warmup), "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": workload_config(c, 1),
            "library_baseline": r, "e2e": r.get("e2e"), "clocks": clocks, "gpu_launches": 0,
            "conv_roofline_frac": round(c["gflop"] * r["value"] / 1e3 / peaks["tensor"], 4)}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------------
# ours
# ---------------------------------------------------------------------------------------------------------------
STAGES = ["enc_l1", "enc_l2", "enc_l3", "enc_l4", "enc_l5", "dec_l1", "dec_l2", "dec_l3", "dec_l4", "dec_l5"]


def run_ours(a):
    import torch
    import torch.distributed as dist

    from xview2_b200 import lib, ops
    from xview2_b200.data_loading.ring import TileRing
    from xview2_b200.model.plt import Model

    c = resolve_config(a)
    train = c["mode"] == "train"
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
    dev = torch.device("cuda", local)
    B, S = c["batch"], c["size"]

    ns = config_namespace(a, c)
    torch.manual_seed(1)
    model = Model(ns).to(dev)
    model = model.train() if train else model.eval()
    opt = flat = None
    if train:
        opt = model.configure_optimizers()
        flat = model.flat
        flat.broadcast_params(0)
        if world > 1 and os.environ.get("XV2_NO_BUCKETS", "0") != "1":
            flat.enable_bucketed_allreduce(model)  # per-stage NCCL all-reduces launched from backward hooks (overlap)

    # synthetic tiles: seeded uint8 "decoded PNG" bytes in the pinned ring; blocky labels
    ring = TileRing(B, S, S, post=c["post"], slots=2, device=dev)
    g = torch.Generator().manual_seed(1 + rank)
    for i in range(ring.slots):
        slot = ring.host(i)
        for k in ("tiles", "tiles_post"):
            if k in slot:
                slot[k].copy_(torch.randint(0, 256, slot[k].shape, generator=g, dtype=torch.uint8))
        cells = torch.randint(0, 5 if c["post"] else 2, (B, S // 32, S // 32), generator=g, dtype=torch.uint8)
        slot["mask"].copy_(cells.repeat_interleave(32, 1).repeat_interleave(32, 2))
    resident = {k: v.to(dev) for k, v in ring.host(0).items()}
    pred_map = torch.empty((B, S, S), dtype=torch.uint8, device=dev) if not train else None

    def train_step(batch):
        opt.zero_grad()
        loss = model.training_step(batch, 0)
        loss.backward()
        n = flat.all_reduce_grads()
        opt.grad_scale = 1.0 / n
        opt.step()
        return loss

    def eval_step(batch):
        """Model.test_step without the .npy dump (plt.py:42-48, 62-67): TTA forward -> argmax label map + F1 counters."""
        with torch.no_grad():
            pred = model.forward(model._image(batch))
            model.f1_score.update(pred, batch["mask"], pred_map)
        return pred_map

    def fence():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    step_fn = train_step if train else eval_step
    eager_step = step_fn
    use_graph = train and os.environ.get("XV2_NO_GRAPH", "0") != "1"
    launch = "eager launches (no CUDA graph)"
    if use_graph:
        # forward + backward (+ bucketed all-reduce + AdamW when the optimizer is capturable) replayed from one CUDA graph
        from xview2_b200.graph import GraphedTrainStep
        try:
            gstep = GraphedTrainStep(model, opt, resident, warmup=2)
            step_fn = gstep  # same signature: batch -> loss tensor
            launch = gstep.describe()
        except Exception as exc:  # noqa: BLE001 -- the capture is an optimisation: the eager step measures the same work
            print(f"[bench] CUDA-graph capture failed ({type(exc).__name__}: {exc}); timing the eager step", file=sys.stderr)
            torch.cuda.synchronize()
            use_graph = False

    # ---- device-resident throughput ("value") ------------------------------------------------------------------
    for _ in range(a.warmup):
        step_fn(resident)
    fence()
    sampler = ClockSampler(local) if rank == 0 else None
    launches0 = lib.launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        result = step_fn(resident)
    e1.record()
    fence()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms)
    launches = lib.launches() - launches0
    last_loss = float(result.detach()) if train else None

    # ---- end to end: pinned host tiles -> side-stream H2D -> step -> result read back, every step ---------------------
    ring.submit(0)
    for i in range(2):  # warm the ring path
        ring.submit(i + 1)
        batch = ring.acquire(i)
        r = step_fn(batch)
        if train:
            float(r.detach())
        ring.release(i)
    fence()
    # The result of EVERY step is read back on the host inside the timed region, one step behind the launch front (the read of
    # step i is issued after step i+1 has been enqueued), so the device never waits for Python between steps.
    if train:
        pinned = torch.empty(a.steps, dtype=torch.float32).pin_memory()
        d2h = 4
    else:
        pinned = torch.empty((2, B, S, S), dtype=torch.uint8).pin_memory()  # the uint8 label maps, double-buffered
        d2h = B * S * S
    host_results = []
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    pending = None
    for j, i in enumerate(range(2, 2 + a.steps)):
        ring.submit(i + 1)           # next batch's copy overlaps this step
        batch = ring.acquire(i)
        r = step_fn(batch)
        ring.release(i)
        if train:
            pinned[j:j + 1].copy_(r.detach().reshape(1), non_blocking=True)  # D2H of this step's result
        else:
            pinned[j & 1].copy_(r, non_blocking=True)
        done = torch.cuda.Event()
        done.record()
        if pending is not None:
            pending[1].synchronize()
            host_results.append(float(pinned[pending[0]]) if train else int(pinned[pending[0] & 1, 0, 0, 0]))
        pending = (j, done)
    pending[1].synchronize()
    host_results.append(float(pinned[pending[0]]) if train else int(pinned[pending[0] & 1, 0, 0, 0]))
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
    peaks = measured_peaks()
    if rank == 0:
        core = model.model
        unets = [m for m in core.modules() if type(m).__name__ == "UNetTemplate"]
        named = []
        for u in (unets or [core]):
            for s in STAGES:
                for suffix in ("", "_pre", "_post"):
                    m = getattr(u, s + suffix, None)
                    if m is not None:
                        named.append((s, m))
        hooks = lib.scope_hooks(named)
        lib.set_scope("fwd:other")
        lib.profile_start()
    if train and rank == 0:
        _orig_backward = torch.Tensor.backward

        def _scoped_backward(self, *args, **kw):  # everything after the forward belongs to the backward scope
            lib.set_scope("bwd")
            return _orig_backward(self, *args, **kw)
        torch.Tensor.backward = _scoped_backward
    try:
        eager_step(resident)
    finally:
        if train and rank == 0:
            torch.Tensor.backward = _orig_backward
    fence()
    if rank == 0:
        rows = lib.profile_stop(per_call=True, with_scope=True)
        for h in hooks:
            h.remove()
        lib.set_scope("")
        prof = {}
        for name, _tag, msv, fl, by, _sc in rows:
            d = prof.setdefault(name, {"calls": 0, "ms": 0.0, "flops": 0.0, "bytes": 0.0})
            d["calls"] += 1
            d["ms"] += msv
            d["flops"] += fl
            d["bytes"] += by
        total = sum(d["ms"] for d in prof.values()) or 1.0
        kernels = {k: {"calls": d["calls"], "ms": round(d["ms"], 3), "share": round(d["ms"] / total, 4),
                       "tflops": round(d["flops"] / d["ms"] / 1e9, 1) if d["flops"] and d["ms"] else None}
                   for k, d in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])}
        stages = {}
        for sc, d in sorted(lib.profile_by_scope(rows).items()):
            stages[sc] = {"ms": round(d["ms"], 3), "launches": d["calls"],
                          "tflops": round(d["flops"] / d["ms"] / 1e9, 1) if d["ms"] else None,
                          "tensor_frac": round(d["flops"] / d["ms"] / 1e9 / peaks["tensor"], 4) if d["ms"] else None,
                          "algorithmic_gbs": round(d["bytes"] / d["ms"] / 1e6, 1) if d["ms"] else None}
        enc = [v for k, v in lib.profile_by_scope(rows).items() if k.startswith("fwd:enc")]
        if enc:
            ems, efl = sum(d["ms"] for d in enc), sum(d["flops"] for d in enc)
            stages["fwd:encoder"] = {"ms": round(ems, 3), "tflops": round(efl / ems / 1e9, 1),
                                     "tensor_frac": round(efl / ems / 1e9 / peaks["tensor"], 4)}
        top = max(prof.items(), key=lambda kv: kv[1]["ms"])
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
                    "note": "the conv entry point runs on the ridge: both the tensor and the HBM fraction are reported",
                    "stages": stages}
        else:
            ach = d["bytes"] / d["ms"] / 1e6  # GB/s
            roof = {"kernel": name, "bound": "hbm", "achieved": round(ach, 1), "peak": peaks["hbm"], "unit": "GB/s",
                    "frac": round(ach / peaks["hbm"], 4), "traffic": traffic, "peak_source": peaks["src"],
                    "launches": d["calls"], "avg_launch_ms": round(d["ms"] / d["calls"], 4), "share_of_step": round(d["ms"] / total, 4),
                    "stages": stages}

    cpu_baseline = cpu_c1 = library = None
    if rank == 0 and world == 1:
        if not a.no_library_baseline:
            del model, opt, flat
            step_fn = eager_step = None
            torch.cuda.empty_cache()
            try:
                library = time_library(c, min(a.steps, 10), 3, dev)
            except Exception as exc:  # noqa: BLE001 -- a reported baseline must not take the measured line down with it
                library = {"unavailable": f"{type(exc).__name__}: {exc}"}
        if not a.no_cpu_baseline:
            v, dt, sample = time_cpu_reference(c, 2, 1)
            cpu_baseline = {"value": v, "unit": c["unit"], "cores": os.cpu_count() or 1, "kind": "port",
                            "sample": sample + "; 1 warm-up + 2 timed steps"}
            if a.config == "c2":
                cpu_c1 = time_cpu_c1()

    if rank == 0:
        value = B * world * a.steps / (ms_total / 1e3)
        line = {
            "metric": c["metric"], "value": value, "unit": c["unit"], "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms_total / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": workload_config(c, world), "launch": launch,
            "e2e": {"value": e2e_value, "unit": c["unit"], "h2d_bytes_per_step": ring.bytes_per_batch, "d2h_bytes_per_step": d2h,
                    "path": ("pinned uint8 HWC tiles + masks -> side-stream H2D (double-buffered TileRing) -> Model.training_step -> backward -> all-reduce -> AdamW -> loss D2H into pinned memory, read on the host every step (one step behind the launch front)"
                             if train else
                             "pinned uint8 HWC tiles + masks -> side-stream H2D (double-buffered TileRing) -> Model.forward (TTA) -> argmax label map + F1 counters -> uint8 label maps D2H into pinned memory, read on the host every step"),
                    "results_read": len(host_results)},
            "gpu_launches": launches,
            # whole-job algorithmic conv FLOP rate against N GPUs' sustained bf16 peak: a per-GPU fraction
            "conv_roofline_frac": round(c["gflop"] * value / 1e3 / (peaks["tensor"] * world), 4),
            "roofline": roof, "cpu_baseline": cpu_baseline, "cpu_baseline_c1": cpu_c1, "library_baseline": library,
            "clocks": clocks, "kernels": kernels, "loss": last_loss,
        }
        json_out.write(json.dumps(line) + "\n")
        json_out.flush()
    if world > 1:
        # The captured CUDA graph holds NCCL work: tearing the process group down underneath it can block in NCCL's watchdog.
        # Every rank has passed the final fence and rank 0 has printed, so leave without the teardown.
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "library"])
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS))
    ap.add_argument("--batch", type=int, default=None, help="units per GPU per step (default: the configuration's)")
    ap.add_argument("--size", type=int, default=1024)
    ap.add_argument("--encoder", default=None)
    ap.add_argument("--no-cpu-baseline", dest="no_cpu_baseline", action="store_true")
    ap.add_argument("--no-library-baseline", dest="no_library_baseline", action="store_true")
    a = ap.parse_args()
    if a.warmup < 3 and a.impl == "ours":
        a.warmup = 3
    if a.impl == "reference":
        run_reference(a)
    elif a.impl == "library":
        run_library(a)
    else:
        run_ours(a)


if __name__ == "__main__":
    main()
