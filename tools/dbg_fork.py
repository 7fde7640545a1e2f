"""GPU debug: reproducibility of a bf16 ResNeSt stage (fork on / off, fused split attention on / off)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from oracle import functional as OF
from xview2_b200 import ops
from xview2_b200.model.encoders import Bottleneck, SplAtBottleneck, _BlockList


def rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-12))


def run(stage, x, gy, fork, fuse):
    ops.FORK_GRADS, ops.FUSE_SPLAT_BN = fork, fuse
    stage.zero_grad(set_to_none=True)
    xi = x.clone().requires_grad_(True)
    y = stage(xi)
    y.backward(gy)
    ops.check_pending_addends()
    torch.cuda.synchronize()
    return y.float().clone(), xi.grad.float().clone(), {k: p.grad.float().clone() for k, p in stage.named_parameters()}


for kind in ("resnest", "resnet"):
    if kind == "resnest":
        blocks = [SplAtBottleneck(128, 64, 1, 1, False, True, 1)] + [SplAtBottleneck(256, 64, 1, 1, False, False, 1) for _ in range(2)]
    else:
        blocks = [Bottleneck(128, 64, 1, 1, True)] + [Bottleneck(256, 64, 1, 1, False) for _ in range(2)]
    stage = _BlockList(blocks)
    shapes = {k: (tuple(v.shape), v.dtype) for k, v in stage.state_dict().items()}
    stage.load_state_dict(OF.deterministic_state(shapes, 11), strict=True)
    stage = stage.cuda().train()
    g = torch.Generator().manual_seed(1)
    x = ops.nhwc(torch.randn(4, 128, 32, 32, generator=g).cuda().to(torch.bfloat16))
    gy = ops.nhwc(torch.randn(4, 256, 32, 32, generator=g).cuda().to(torch.bfloat16))
    ref = run(stage, x, gy, False, True)
    for name, fork, fuse in (("nofork again", False, True), ("fork", True, True), ("nofork unfused", False, False), ("fork unfused", True, False)):
        r = run(stage, x, gy, fork, fuse)
        worst = max(((rel(r[2][k], ref[2][k]), k) for k in ref[2]))
        print(f"{kind:8s} {name:16s} y {rel(r[0], ref[0]):.2e}  gx {rel(r[1], ref[1]):.2e}  worst param grad {worst[0]:.2e} ({worst[1]})", flush=True)
