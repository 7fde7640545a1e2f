"""Flat-buffer parameters, fused optimizers and the data-parallel gradient all-reduce.

The reference selects ``torch.optim.AdamW`` / apex ``FusedAdam`` / ``FusedSGD`` in Model.configure_optimizers
(/root/reference/model/plt.py:150-160) and leaves the gradient all-reduce to DDP (main.py:106-107).  Here every
parameter of the model is re-homed as a view into ONE fp32 buffer (and its gradient into a second one), so that

  * the optimizer step is a single libxv2 launch over the whole buffer (xv2_adamw), not one launch per tensor;
  * the data-parallel exchange is a single NCCL all-reduce(SUM) over the flat gradient buffer followed by 1/N folded
    into the optimizer's gradient scale -- the only collective of the path;
  * zeroing gradients is one memset.
"""
import torch

from . import lib, ops


class FlatParams:
    """Re-homes ``module``'s parameters (and gradients) into two flat fp32 buffers, preserving each tensor's physical
    (channels-last) element order so kernels keep seeing [K][R][S][C] weights."""

    def __init__(self, module):
        params, seen = [], set()
        for p in module.parameters():
            if id(p) not in seen:
                seen.add(id(p))
                params.append(p)
        if not params:
            raise ValueError("module has no parameters")
        dev = params[0].device
        total = sum((p.numel() + 3) // 4 * 4 for p in params)  # 16-byte aligned slots
        self.data = torch.zeros(total, dtype=torch.float32, device=dev)
        self.grad = torch.zeros(total, dtype=torch.float32, device=dev)
        self.params = params
        off = 0
        for p in params:
            n = p.numel()
            if p.dtype != torch.float32:
                raise lib.Xv2Error("master parameters must be fp32")
            if p.dim() == 4:
                a, b, r, s = p.shape
                phys = ops.nhwc(p.data)  # logical (a, b, r, s), physical [a][r][s][b]
                view = self.data[off:off + n].view(a, r, s, b).permute(0, 3, 1, 2)
                gview = self.grad[off:off + n].view(a, r, s, b).permute(0, 3, 1, 2)
                view.copy_(phys)
            else:
                view = self.data[off:off + n].view(p.shape)
                gview = self.grad[off:off + n].view(p.shape)
                view.copy_(p.data)
            p.data = view
            p.grad = gview
            p._xv2_flat = True  # kernels may accumulate weight gradients straight into p.grad (ops._grad_buffer)
            off += (n + 3) // 4 * 4
        self.numel = total
        self._offsets = {}
        off = 0
        for p in params:
            self._offsets[id(p)] = (off, off + p.numel())
            off += (p.numel() + 3) // 4 * 4
        self._buckets = []      # [{"lo", "hi", "visits", "seen", "done"}] in flat-buffer order (bucketed all-reduce)
        self._works = []
        self._hooks = []
        ops.clear_weight_cache()

    # -- bucketed gradient all-reduce, overlapped with the backward pass ---------------------------------------------------
    def enable_bucketed_allreduce(self, module, min_bytes=8 << 20, group=None):
        """DDP-style overlap (main.py:106-107) on the flat buffer: the parameters of each network stage (enc_l*, dec_l*, fusion
        convs) are one contiguous range; a full-backward hook on the stage launches the all-reduce of that range as soon as the
        stage's backward has run (async on NCCL's own stream, so it overlaps the backward of the stages below).  Whatever is
        left (the stem, heads, stages whose hook cannot fire) is reduced by all_reduce_grads() at the end of backward, which also
        waits for every outstanding collective.  Adjacent ranges are merged up to `min_bytes` so no collective is latency-bound."""
        self.disable_bucketed_allreduce()
        self._group = group
        stages = []
        seen_mods = set()
        for name, mod in module.named_modules():
            leaf = name.rsplit(".", 1)[-1]
            is_stage = (leaf.startswith(("enc_l", "dec_l")) and not leaf.startswith("enc_l1")) or leaf in ("conv_pre", "conv_post")
            if not is_stage or id(mod) in seen_mods or any(name.startswith(t[0] + ".") for t in stages):
                continue
            seen_mods.add(id(mod))
            rng = [self._offsets[id(p)] for p in mod.parameters() if id(p) in self._offsets]
            if rng:
                stages.append((name, mod, min(r[0] for r in rng), max(r[1] for r in rng)))
        # keep only stages whose ranges are disjoint and ordered (shared modules registered twice keep their first spelling)
        stages.sort(key=lambda t: t[2])
        buckets, cur = [], None
        for name, mod, lo, hi in stages:
            if cur is not None and lo < cur["hi"]:
                continue
            if cur is not None and (cur["hi"] - cur["lo"]) * 4 < min_bytes and lo == cur["hi"]:
                cur["hi"] = hi
                cur["mods"].append(mod)
                continue
            cur = {"lo": lo, "hi": hi, "mods": [mod], "visits": 0, "seen": 0, "done": False}
            buckets.append(cur)
        self._buckets = buckets
        for b in buckets:
            for mod in b["mods"]:
                self._hooks.append(mod.register_forward_pre_hook(lambda _m, _i, b=b: self._count_visit(b)))
                self._hooks.append(mod.register_full_backward_hook(lambda _m, _gi, _go, b=b: self._stage_done(b)))
        return len(buckets)

    def disable_bucketed_allreduce(self):
        for h in self._hooks:
            h.remove()
        self._hooks, self._buckets, self._works = [], [], []

    def _dist_world(self):
        dist = torch.distributed
        if not (dist.is_available() and dist.is_initialized()):
            return 1
        return dist.get_world_size(getattr(self, "_group", None))

    def _count_visit(self, b):
        if torch.is_grad_enabled():
            b["visits"] += 1

    def _stage_done(self, b):
        b["seen"] += 1
        if b["done"] or b["seen"] < b["visits"] or self._dist_world() <= 1:
            return
        b["done"] = True
        ops.sync_side_streams()
        self._works.append(torch.distributed.all_reduce(self.grad[b["lo"]:b["hi"]], group=getattr(self, "_group", None), async_op=True))

    def _reset_buckets(self):
        for b in self._buckets:
            b["visits"] = b["seen"] = 0
            b["done"] = False

    def zero_grad(self):
        ops.sync_side_streams()
        self._reset_buckets()
        self.grad.zero_()
        if self.grad.is_cuda:
            ops.new_step_scratch(self.grad.device)  # the step's zero-initialised scratch arena: one memset

    def all_reduce_grads(self, group=None):
        """all-reduce(SUM) of the flat gradient buffer over the data-parallel ranks; returns the world size (the 1/N
        average is applied by the optimizer through its ``grad_scale``)."""
        ops.sync_side_streams()  # weight gradients issued on the side stream have landed in the flat buffer
        ops.check_pending_addends()  # backward is over: a gradient part still parked by ops.fork would be a silently wrong step
        dist = torch.distributed
        if not (dist.is_available() and dist.is_initialized()):
            return 1
        world = dist.get_world_size(group)
        if world > 1:
            if not self._buckets:
                dist.all_reduce(self.grad, group=group)
            else:  # what the stage hooks have not launched yet: the complement of the finished ranges, as contiguous pieces
                pos = 0
                for b in sorted(self._buckets, key=lambda b: b["lo"]):
                    if b["done"]:
                        if b["lo"] > pos:
                            self._works.append(dist.all_reduce(self.grad[pos:b["lo"]], group=group, async_op=True))
                        pos = b["hi"]
                if pos < self.numel:
                    self._works.append(dist.all_reduce(self.grad[pos:self.numel], group=group, async_op=True))
                for w in self._works:
                    w.wait()  # the current stream waits for NCCL's stream (an event wait: capturable, no host block)
                self._works = []
        return world

    def broadcast_params(self, src=0, group=None):
        dist = torch.distributed
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.broadcast(self.data, src, group=group)
            ops.clear_weight_cache()


class _FlatOptimizer:
    """Minimal torch.optim-like surface (param_groups / step / zero_grad / state_dict) over a FlatParams."""

    def __init__(self, flat, **defaults):
        self.flat = flat
        self.param_groups = [dict(params=flat.params, **defaults)]
        self.step_count = 0
        self.grad_scale = 1.0

    def zero_grad(self, set_to_none=False):
        self.flat.zero_grad()

    def state_dict(self):
        return {"step": self.step_count, "param_groups": [{k: v for k, v in g.items() if k != "params"} for g in self.param_groups],
                "state": {k: v for k, v in self._state().items()}}

    def load_state_dict(self, sd):
        self.step_count = sd["step"]
        for g, s in zip(self.param_groups, sd["param_groups"]):
            g.update(s)
        for k, v in sd["state"].items():
            getattr(self, k).copy_(v)

    def _state(self):
        return {}

    # -- CUDA-graph-capturable step: the per-step scalars live in device memory --------------------------------------------
    def _hyper_values(self):
        raise NotImplementedError

    def prepare_step(self):
        """Host side of a captured step: advances the step count and refreshes the device block of per-step scalars (learning
        rate, bias corrections, gradient scale) with ONE async 32-byte copy on the current stream."""
        self.step_count += 1
        vals = torch.tensor(self._hyper_values(), dtype=torch.float32).pin_memory()
        if getattr(self, "hyper", None) is None:
            self.hyper = torch.zeros(8, dtype=torch.float32, device=self.flat.data.device)
        self.hyper[:vals.numel()].copy_(vals, non_blocking=True)

    def step_captured(self):
        """Device side of a captured step (no host scalar in the launch arguments): optimizer update + batched weight re-pack."""
        raise NotImplementedError



class FusedAdamW(_FlatOptimizer):
    """torch.optim.AdamW semantics (decoupled weight decay, bias correction) in one launch over the flat buffer.
    Replaces plt.py:154 ('adamw') and apex FusedAdam (plt.py:153: adam_w_mode defaults to True, i.e. also AdamW)."""

    def __init__(self, flat, lr=3e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        super().__init__(flat, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        self.exp_avg = torch.zeros_like(flat.data)
        self.exp_avg_sq = torch.zeros_like(flat.data)

    def _state(self):
        return {"exp_avg": self.exp_avg, "exp_avg_sq": self.exp_avg_sq}

    def step(self):
        ops.sync_side_streams()
        g = self.param_groups[0]
        self.step_count += 1
        ops.adamw_step(self.flat.data, self.flat.grad, self.exp_avg, self.exp_avg_sq, g["lr"], g["betas"][0], g["betas"][1],
                       g["eps"], g["weight_decay"], self.step_count, self.grad_scale)
        ops.clear_weight_cache()  # packed bf16 weight copies are stale now ...
        ops.repack_all()          # ... and refreshed by one batched launch

    def _hyper_values(self):
        import math
        g = self.param_groups[0]
        b1, b2 = g["betas"]
        t = self.step_count
        return [g["lr"], b1, b2, g["eps"], g["weight_decay"], 1.0 - b1 ** t, math.sqrt(1.0 - b2 ** t), self.grad_scale]

    def step_captured(self):
        ops.sync_side_streams()
        lib.call("xv2_adamw_dev", self.flat.data.data_ptr(), self.flat.grad.data_ptr(), self.exp_avg.data_ptr(),
                 self.exp_avg_sq.data_ptr(), self.flat.data.numel(), self.hyper.data_ptr())
        ops.clear_weight_cache()
        ops.repack_all()


class FusedSGD(_FlatOptimizer):
    """SGD with momentum (apex FusedSGD defaults: dampening 0, no nesterov, no weight decay), plt.py:152."""

    def __init__(self, flat, lr=3e-4, momentum=0.9):
        super().__init__(flat, lr=lr, momentum=momentum)
        self.momentum_buffer = torch.zeros_like(flat.data)

    def _state(self):
        return {"momentum_buffer": self.momentum_buffer}

    def step(self):
        ops.sync_side_streams()
        g = self.param_groups[0]
        self.step_count += 1
        ops.sgd_step(self.flat.data, self.flat.grad, self.momentum_buffer, g["lr"], g["momentum"], self.grad_scale,
                     self.step_count)
        ops.clear_weight_cache()
        ops.repack_all()

    def _hyper_values(self):
        g = self.param_groups[0]
        return [g["lr"], g["momentum"], self.grad_scale, 1.0 if self.step_count == 1 else 0.0]

    def step_captured(self):
        ops.sync_side_streams()
        lib.call("xv2_sgd_dev", self.flat.data.data_ptr(), self.flat.grad.data_ptr(), self.momentum_buffer.data_ptr(),
                 self.flat.data.numel(), self.hyper.data_ptr())
        ops.clear_weight_cache()
        ops.repack_all()
