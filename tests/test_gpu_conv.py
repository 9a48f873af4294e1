"""tcgen05 implicit-GEMM convolution kernels vs a plain PyTorch fp32 reference of the same op
(inputs rounded to bf16 on both sides; the kernel accumulates in fp32 and stores bf16)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    import hgb200
    return hgb200.ops


@pytest.fixture(scope="module")
def torch():
    import torch
    return torch


def _ref_conv(torch, x, w, bias, ksize, relu, flip=False):
    """x (N,H,W,Cin) bf16, w (Cout, k*k*Cin) bf16 [tap-major] -> fp32 NHWC."""
    import torch.nn.functional as F
    N, H, W, Cin = x.shape
    Cout = w.shape[0]
    w4 = w.float().reshape(Cout, ksize, ksize, Cin).permute(0, 3, 1, 2)          # OIHW
    if flip:
        w4 = w4.flip(2, 3)
    y = F.conv2d(x.float().permute(0, 3, 1, 2), w4, bias, padding=ksize // 2)
    y = y.permute(0, 2, 3, 1)
    return torch.relu(y) if relu else y


def _rand(torch, shape, seed, scale=1.0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(shape, device="cuda", generator=g) * scale).to(torch.bfloat16)


CASES = [
    # N, H, Cin, Cout, k
    (2, 64, 256, 128, 1),
    (2, 64, 128, 128, 3),
    (2, 64, 128, 256, 1),
    (3, 32, 128, 128, 3),
    (2, 16, 256, 256, 1),
    (5, 8, 128, 128, 3),
    (3, 4, 128, 128, 3),
    (1, 4, 256, 128, 1),
    (1, 128, 64, 64, 3),
    (1, 128, 64, 128, 1),
    (2, 64, 64, 256, 1),
    (1, 64, 192, 64, 1),
    # >= 8*148 pixel tiles: the weight-stationary 4-tile path of the 3x3 kernel (incl. a group tail: 1186 tiles)
    (40, 64, 128, 128, 3),
    (593, 16, 128, 128, 3),
    (152, 32, 128, 128, 3),      # strip-reuse (HALO) path at 32x32: four image rows per tile
    (10, 128, 64, 64, 3),
]


@pytest.mark.parametrize("N,H,Cin,Cout,k", CASES)
def test_conv_forward(ops, torch, N, H, Cin, Cout, k):
    x = _rand(torch, (N, H, H, Cin), 1)
    w = _rand(torch, (Cout, k * k * Cin), 2, scale=(k * k * Cin) ** -0.5)
    bias = torch.randn(Cout, device="cuda") * 0.1
    stats = torch.zeros(2 * Cout, device="cuda")
    y = ops.conv_gemm(x, w, bias=bias, ksize=k, relu=True, stats=stats)
    ref = _ref_conv(torch, x, w, bias, k, True)
    torch.cuda.synchronize()
    err = (y.float() - ref).abs().max().item()
    assert err <= 2e-2 * max(1.0, ref.abs().max().item()), f"max err {err}"      # bf16 output rounding
    yf = y.float().reshape(-1, Cout)
    np.testing.assert_allclose(stats[:Cout].cpu().numpy(), yf.sum(0).cpu().numpy(), rtol=1e-3, atol=1e-2)
    np.testing.assert_allclose(stats[Cout:].cpu().numpy(), (yf * yf).sum(0).cpu().numpy(), rtol=1e-3, atol=1e-2)


def test_conv_linear_residuals_and_pitch(ops, torch):
    N, H, Cin, Cout = 2, 64, 256, 256
    x = _rand(torch, (N, H, H, Cin), 3)
    w = _rand(torch, (Cout, Cin), 4, scale=Cin ** -0.5)
    r1 = _rand(torch, (N, H, H, Cout), 5)
    r2 = _rand(torch, (N, H, H, Cout), 6)
    y = ops.conv_gemm(x, w, res1=r1, res2=r2, ksize=1, relu=False)
    ref = _ref_conv(torch, x, w, None, 1, False) + r1.float() + r2.float()
    assert (y.float() - ref).abs().max().item() <= 3e-2 * max(1.0, ref.abs().max().item())
    # in-place accumulate: res1 is the output buffer itself
    y2 = r1.clone()
    ops.conv_gemm(x, w, res1=y2, ksize=1, relu=False, out=y2)
    ref2 = _ref_conv(torch, x, w, None, 1, False) + r1.float()
    assert (y2.float() - ref2).abs().max().item() <= 3e-2 * max(1.0, ref2.abs().max().item())


@pytest.mark.parametrize("N,H,C", [(2, 64, 128), (3, 8, 128), (2, 4, 128), (38, 64, 128), (150, 32, 128)])
def test_conv_dgrad_mirrored_taps(ops, torch, N, H, C):
    """tap_sign=-1 with the same tap-major weights == correlation with the flipped kernel."""
    dy = _rand(torch, (N, H, H, C), 7)
    w = _rand(torch, (C, 9 * C), 8, scale=(9 * C) ** -0.5)
    res = _rand(torch, (N, H, H, C), 11)
    y = ops.conv_gemm(dy, w, ksize=3, relu=False, tap_sign=-1, res1=res)
    ref = _ref_conv(torch, dy, w, None, 3, False, flip=True) + res.float()
    assert (y.float() - ref).abs().max().item() <= 2e-2 * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize("N,H,tap_sign", [(40, 64, 1), (152, 32, 1), (38, 64, -1), (150, 32, -1)])
def test_conv3x3_cta_pairs_opt_in(ops, torch, N, H, tap_sign):
    """The cta_group::2 variant of the strip-reuse 3x3 kernel (hgb_debug_set(30, 1): two CTAs of one TPC, M = 256 MMAs,
    half of every weight box per CTA) against fp32 torch, forward and mirrored-tap dgrad with a residual, and bit-identical
    to the default one-CTA kernel (same MMA order per accumulator, same epilogue)."""
    import hgb200
    lib = hgb200._lib.lib
    C = 128
    x = _rand(torch, (N, H, H, C), 21)
    w = _rand(torch, (C, 9 * C), 22, scale=(9 * C) ** -0.5)
    bias = torch.randn(C, device="cuda") * 0.1 if tap_sign == 1 else None
    res = _rand(torch, (N, H, H, C), 23) if tap_sign == -1 else None
    stats1, stats2 = torch.zeros(2 * C, device="cuda"), torch.zeros(2 * C, device="cuda")
    y_one = ops.conv_gemm(x, w, bias=bias, res1=res, ksize=3, relu=tap_sign == 1, tap_sign=tap_sign, stats=stats1)
    try:
        lib.hgb_debug_set(30, 1)
        y_pair = ops.conv_gemm(x, w, bias=bias, res1=res, ksize=3, relu=tap_sign == 1, tap_sign=tap_sign, stats=stats2)
        torch.cuda.synchronize()
    finally:
        lib.hgb_debug_set(30, 0)
    ref = _ref_conv(torch, x, w, bias, 3, tap_sign == 1, flip=tap_sign == -1)
    if res is not None:
        ref = ref + res.float()
    assert (y_pair.float() - ref).abs().max().item() <= 2e-2 * max(1.0, ref.abs().max().item())
    assert torch.equal(y_pair, y_one)
    np.testing.assert_allclose(stats2.cpu().numpy(), stats1.cpu().numpy(), rtol=1e-4, atol=1e-2)


WG_CASES = [
    (2, 64, 128, 128, 3),
    (2, 64, 256, 128, 1),
    (2, 64, 128, 256, 1),
    (3, 16, 256, 256, 1),
    (5, 8, 128, 128, 3),
    (3, 4, 128, 128, 3),
    (1, 128, 64, 64, 3),
    (1, 64, 192, 64, 1),
    (2, 64, 64, 256, 1),
    # >= 8 * 148 pixel tiles: the wide CTA tiles of the HBM-bound 1x1 layers (<256,1>, <128,2>) and their neighbours
    (40, 64, 256, 128, 1),
    (40, 64, 128, 256, 1),
    (38, 64, 256, 256, 1),
    (40, 64, 256, 64, 1),
    # >= 4 * 148 pixel tiles: the rolling-strip 3x3 kernel (three vertical taps per CTA; ranges start mid-image)
    (20, 64, 128, 128, 3),
    (19, 64, 128, 128, 3),
    (75, 32, 128, 128, 3),
]


@pytest.mark.parametrize("N,H,Cin,Cout,k", WG_CASES)
def test_conv_wgrad(ops, torch, N, H, Cin, Cout, k):
    import torch.nn.functional as F
    x = _rand(torch, (N, H, H, Cin), 9)
    dy = _rand(torch, (N, H, H, Cout), 10)
    dw = ops.conv_wgrad(x, dy, ksize=k)
    xr = x.float().permute(0, 3, 1, 2).requires_grad_(False)
    wr = torch.zeros((Cout, Cin, k, k), device="cuda", requires_grad=True)
    F.conv2d(xr, wr, padding=k // 2).backward(dy.float().permute(0, 3, 1, 2))
    ref = wr.grad.permute(0, 2, 3, 1).reshape(Cout, k * k * Cin)                 # tap-major, channel-minor
    scale = ref.abs().max().item()
    err = (dw - ref).abs().max().item()
    assert err <= 2e-3 * max(1.0, scale), f"max err {err} (scale {scale})"      # fp32 accumulate, atomics order
