"""GPU tests (-m gpu) of the tcgen05 implicit-GEMM convolution kernels (forward, dgrad, wgrad, transposed conv)."""
import pytest

from tests.tc_cases import CASES, run_case

pytestmark = pytest.mark.gpu


from tests.tc_cases import run_f32_case  # noqa: E402


@pytest.mark.parametrize("name", list(CASES))
def test_tc_conv(name):
    """bf16-output kernels: every element within ONE bf16 ulp of torch's fp32 result on the same bf16 operands (+ 5e-4 of the
    largest magnitude for fp32 accumulation-order noise); the fp32 weight gradient within 1e-3.  A kernel that drops a K block,
    a tap or a halo column fails these by orders of magnitude (the former 2e-2 bound would not have caught a deep-layer drop)."""
    errs = run_case(name)
    for key, val in errs.items():
        if key.endswith("_ulp"):
            assert val < 5e-4, (name, key, errs)
        elif key == "wgrad":
            assert val < 1e-3, (name, key, errs)
        else:
            assert val < 8e-3, (name, key, errs)  # one bf16 ulp (2^-8 = 3.9e-3) of the largest element, plus slack


@pytest.mark.parametrize("name", [n for n in CASES if not n.startswith("strip_")])
def test_tc_conv_f32_output(name):
    """fp32-output epilogue of the same tcgen05 main loop (north_star: 1e-3 on fp32 logits): <= 1e-3, expected ~1e-5."""
    err = run_f32_case(name)
    assert err < 1e-3, (name, err)


@pytest.mark.parametrize("shape", [(2, 32, 0, 64, 256, 32, 1, 3), (2, 64, 64, 24, 128, 64, 1, 3), (1, 128, 0, 40, 128, 256, 2, 3),
                                   (2, 64, 0, 32, 32, 256, 1, 1), (2, 256, 0, 16, 16, 96, 1, 3), (3, 128, 64, 24, 64, 512, 1, 3),
                                   (2, 128, 0, 16, 16, 256, 2, 3), (2, 512, 0, 8, 16, 2048, 1, 1)])
def test_conv_stats_epilogue(shape):
    """BN statistics fused into the conv epilogues (strip and tile kernels) == sums over the stored (bf16-rounded) outputs."""
    import torch

    from xview2_b200 import ops
    n, c0, c1, h, w, k, groups, r = shape
    g = torch.Generator().manual_seed(5)
    cl = torch.channels_last
    x = torch.randn(n, c0, h, w, generator=g).cuda().to(torch.bfloat16).contiguous(memory_format=cl)
    x2 = torch.randn(n, c1, h, w, generator=g).cuda().to(torch.bfloat16).contiguous(memory_format=cl) if c1 else None
    cg = (c0 + c1) // groups
    wt = (torch.randn(k, cg, r, r, generator=g) * (2.0 / (r * r * cg)) ** 0.5).cuda().contiguous(memory_format=cl)
    out, stats = ops.conv2d_stats(x, wt, None, 1, r // 2, 1, groups, x2)
    assert stats is not None, "tensor-core kernels declined the fused statistics"
    o = out.double()
    ref = torch.cat((o.sum((0, 2, 3)), (o * o).sum((0, 2, 3))))
    import torch.nn.functional as F
    src = x.float() if x2 is None else torch.cat((x.float(), x2.float()), 1)
    yr = F.conv2d(src, wt.to(torch.bfloat16).float(), None, 1, r // 2, 1, groups)
    from tests.tc_cases import ulp_excess
    assert ulp_excess(out, yr) < 5e-4
    err = float((stats - ref).abs().max() / ref.abs().max())
    assert err < 1e-5, err
