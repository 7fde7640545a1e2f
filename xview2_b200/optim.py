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
        ops.clear_weight_cache()

    def zero_grad(self):
        ops.sync_side_streams()
        self.grad.zero_()
        if self.grad.is_cuda:
            ops.new_step_scratch(self.grad.device)  # the step's zero-initialised scratch arena: one memset

    def all_reduce_grads(self, group=None):
        """all-reduce(SUM) of the flat gradient buffer over the data-parallel ranks; returns the world size (the 1/N
        average is applied by the optimizer through its ``grad_scale``)."""
        ops.sync_side_streams()  # weight gradients issued on the side stream have landed in the flat buffer
        dist = torch.distributed
        if not (dist.is_available() and dist.is_initialized()):
            return 1
        world = dist.get_world_size(group)
        if world > 1:
            dist.all_reduce(self.grad, group=group)
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
