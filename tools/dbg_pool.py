import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from xview2_b200 import ops
CL = torch.channels_last
torch.manual_seed(0)
for dtype in (torch.float32,):
    x = torch.randn(2, 16, 15, 18).to(dtype).cuda().contiguous(memory_format=CL).requires_grad_(True)
    cases = [("max321", lambda a: ops.max_pool2d(a, 3, 2, 1), lambda a: F.max_pool2d(a, 3, 2, 1)),
             ("avg321", lambda a: ops.avg_pool2d(a, 3, 2, 1), lambda a: F.avg_pool2d(a, 3, 2, 1)),
             ("avg311", lambda a: ops.avg_pool2d(a, 3, 1, 1), lambda a: F.avg_pool2d(a, 3, 1, 1)),
             ("avg220c", lambda a: ops.avg_pool2d(a, 2, 2, 0, True, False), lambda a: F.avg_pool2d(a, 2, 2, 0, True, False))]
    for name, fn, rf in cases:
        x.grad = None
        y = fn(x)
        gy = torch.randn(*y.shape, device="cuda").to(dtype)
        y.backward(gy)
        ours = x.grad.detach().float().cpu()
        xc = x.detach().float().cpu().contiguous().requires_grad_(True)
        yc = rf(xc)
        yc.backward(gy.float().cpu())
        xg = x.detach().float().clone().requires_grad_(True)
        yg = rf(xg)
        yg.backward(gy.float())
        xn = x.detach().float().contiguous().requires_grad_(True)
        yn = rf(xn); yn.backward(gy.float().contiguous())
        print(name, "fwd vs cpu", float((y.float().cpu() - yc).abs().max()), "bwd ours vs cpu", float((ours - xc.grad).abs().max()),
              "torch cuda CL vs cpu", float((xg.grad.cpu() - xc.grad).abs().max()), "torch cuda NCHW vs cpu", float((xn.grad.cpu() - xc.grad).abs().max()))
