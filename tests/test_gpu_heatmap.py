"""Parity of the CUDA heatmap kernels (through the C ABI) with the oracle and the golden
fixtures produced by the reference.  Bit-exact for indices / coordinates / targets."""
import os

import numpy as np
import pytest

from oracle import heatmap_oracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    import hgb200
    return hgb200.ops


@pytest.fixture(scope="module")
def torch():
    import torch
    return torch


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


# ---------------------------------------------------------------- decode
def test_decode_golden_f32(ops, golden_dir):
    g = _load(golden_dir, "decode_golden.npz")
    hm = g["heatmaps"]
    for ti, thr in enumerate(g["thresholds"]):
        for ver in (1, 2):
            idx, kp = ops.decode_batch(hm, float(thr), ver)
            ref = g[f"v{ver}_thr{ti}"]
            np.testing.assert_array_equal(kp.cpu().numpy().view(np.uint32), ref.view(np.uint32))
            oidx, _ = orc.decode_batch(hm, float(thr), ver)
            np.testing.assert_array_equal(idx.cpu().numpy(), oidx)


def test_decode_does_not_modify_input(ops, torch, golden_dir):
    g = _load(golden_dir, "decode_golden.npz")
    hm = torch.as_tensor(g["heatmaps"], device="cuda")
    before = hm.clone()
    ops.decode_batch(hm, 1e-6, 2)
    assert torch.equal(hm.view(torch.int32), before.view(torch.int32))


@pytest.mark.parametrize("size,batch", [(64, 7), (128, 3), (8, 5), (2, 4)])
def test_decode_random_vs_oracle(ops, size, batch):
    rng = np.random.default_rng(size * 100 + batch)
    # quantised values -> many exact ties, exercising the lowest-index rule
    hm = (rng.integers(0, 50, size=(batch, size, size, 17)) / 50.0).astype(np.float32)
    for ver in (1, 2):
        idx, kp = ops.decode_batch(hm, 0.3, ver)
        oidx, okp = orc.decode_batch(hm, 0.3, ver)
        np.testing.assert_array_equal(idx.cpu().numpy(), oidx)
        np.testing.assert_array_equal(kp.cpu().numpy().view(np.uint32), okp.view(np.uint32))


def test_decode_bf16(ops, torch):
    rng = np.random.default_rng(5)
    hm = torch.as_tensor(rng.random((6, 64, 64, 17), dtype=np.float32), device="cuda").to(torch.bfloat16)
    idx, kp = ops.decode_batch(hm, 1e-6, 2)
    oidx, okp = orc.decode_batch(hm.float().cpu().numpy(), 1e-6, 2)
    np.testing.assert_array_equal(idx.cpu().numpy(), oidx)
    np.testing.assert_array_equal(kp.cpu().numpy().view(np.uint32), okp.view(np.uint32))


def test_decode_large_batch_property(ops, torch):
    """Full-size case (B=1024): the decoded index must point at a value equal to the map maximum
    and no earlier element may equal it (size-independent property, no oracle loop)."""
    B = 1024
    gen = torch.Generator(device="cuda").manual_seed(3)
    hm = torch.rand((B, 64, 64, 17), device="cuda", generator=gen)
    idx, kp = ops.decode_batch(hm, 1e-6, 1)
    flat = hm.reshape(B, 4096, 17)
    mx = flat.max(dim=1).values
    first = ((flat == mx[:, None, :]).int().cumsum(1) == 0).sum(1)          # index of the first maximum
    assert torch.equal(idx[:, :, 0].long(), first)
    assert torch.equal(kp[:, :, 2], mx)
    assert torch.equal(idx[:, :, 1], idx[:, :, 0] % 64) and torch.equal(idx[:, :, 2], idx[:, :, 0] // 64)


def test_decode_rejects_bad_shapes(ops, torch):
    with pytest.raises(ValueError):
        ops.decode_batch(torch.zeros((1, 32, 64, 17), device="cuda"))
    with pytest.raises(ValueError):
        ops.decode_batch(torch.zeros((1, 64, 64, 17), device="cuda"), version=3)
    idx, kp = ops.decode_batch(torch.zeros((0, 64, 64, 17), device="cuda"))
    assert idx.shape == (0, 17, 4) and kp.shape == (0, 17, 3)


# ---------------------------------------------------------------- render
def test_render_golden(ops, golden_dir):
    g = _load(golden_dir, "render_golden.npz")
    out = ops.render_targets(g["kps_x"], g["kps_y"], g["kps_v"], 64, 64).cpu().numpy()
    np.testing.assert_array_equal(out.view(np.uint32), g["targets"].view(np.uint32))
    out128 = ops.render_targets(g["kps_x128"], g["kps_y128"], g["kps_v"][:3], 128, 128).cpu().numpy()
    ref128 = np.zeros((3, 128, 128, 17), np.float32)
    ref128[tuple(g["t128_idx"])] = g["t128_val"]
    np.testing.assert_array_equal(out128.view(np.uint32), ref128.view(np.uint32))


@pytest.mark.parametrize("size,batch", [(64, 33), (128, 9)])
def test_render_random_vs_oracle(ops, size, batch):
    rng = np.random.default_rng(size + batch)
    kx = rng.uniform(-4, size + 4, (batch, 17)).astype(np.float32)
    ky = rng.uniform(-4, size + 4, (batch, 17)).astype(np.float32)
    kv = rng.choice([0, 1, 2], p=[.2, .3, .5], size=(batch, 17))
    out = ops.render_targets(kx, ky, kv, size, size).cpu().numpy()
    ref = orc.render_targets(kx, ky, kv, size, size)
    np.testing.assert_array_equal(out.view(np.uint32), ref.view(np.uint32))


def test_render_large_batch_property(ops, torch):
    """B=4096: every drawn joint contributes sum(G) exactly when its 7x7 stamp is interior."""
    B = 4096
    rng = np.random.default_rng(9)
    kx = rng.uniform(4, 59, (B, 17)).astype(np.float32)
    ky = rng.uniform(4, 59, (B, 17)).astype(np.float32)
    kv = rng.integers(0, 3, (B, 17))
    out = ops.render_targets(kx, ky, kv, 64, 64)
    sums = out.sum(dim=(1, 2)).cpu().numpy()
    gsum = float(orc.gaussian_patch(1).astype(np.float32).sum(dtype=np.float64))
    np.testing.assert_allclose(sums, (kv > 0) * gsum, rtol=1e-5, atol=1e-5)
    assert int((out == 1.0).sum().item()) == int((kv > 0).sum())


# ---------------------------------------------------------------- losses
def _loss_case(seed, B=5, size=64):
    rng = np.random.default_rng(seed)
    kx = rng.uniform(-4, size + 4, (B, 17)).astype(np.float32)
    ky = rng.uniform(-4, size + 4, (B, 17)).astype(np.float32)
    kv = rng.choice([0, 1, 2], p=[.2, .3, .5], size=(B, 17))
    t = orc.render_targets(kx, ky, kv, size, size)
    p = (1.0 / (1.0 + np.exp(-rng.standard_normal(t.shape)))).astype(np.float32)
    return t, p


@pytest.mark.parametrize("kind", ["weighted_mse", "mse", "iou", "weighted_keypoint_mse"])
def test_loss_and_grad_f32(ops, kind):
    t, p = _loss_case(11)
    loss, grad = ops.loss_fwd_bwd(kind, t, p)
    oloss, ograd = orc.loss_and_grad(kind, t, p)
    assert abs(loss.item() - oloss) <= 1e-6 * max(1.0, abs(oloss))     # tolerance: fp32 products, fp64 sums
    g = grad.cpu().numpy().astype(np.float64)
    np.testing.assert_allclose(g, ograd, rtol=2e-5, atol=1e-12)


@pytest.mark.parametrize("kind", ["weighted_mse", "iou"])
def test_loss_bf16_pred(ops, torch, kind):
    t, p = _loss_case(12)
    pb = torch.as_tensor(p, device="cuda").to(torch.bfloat16)
    loss, grad = ops.loss_fwd_bwd(kind, t, pb)
    assert grad.dtype == torch.bfloat16
    oloss, ograd = orc.loss_and_grad(kind, t, pb.float().cpu().numpy())
    assert abs(loss.item() - oloss) <= 1e-6 * max(1.0, abs(oloss))
    np.testing.assert_allclose(grad.float().cpu().numpy(), ograd, rtol=1e-2, atol=1e-10)   # bf16 output rounding


@pytest.mark.parametrize("kind,fn", [("weighted_mse", orc.weighted_mse_map), ("mse", orc.mse_map),
                                     ("weighted_keypoint_mse", orc.keypoint_mse_map), ("iou", orc.iou_vec)])
def test_loss_maps(ops, kind, fn):
    t, p = _loss_case(13, B=3)
    out = ops.loss_map(kind, t, p).cpu().numpy()
    ref = fn(t, p)
    assert out.shape == ref.shape
    np.testing.assert_allclose(out, ref, rtol=1e-5, atol=1e-8)


def test_loss_global_batch_scaling(ops):
    """A shard's loss/grad with global_batch=2B is exactly half the stand-alone value (DP contract)."""
    t, p = _loss_case(14, B=4)
    l1, g1 = ops.loss_fwd_bwd("weighted_mse", t, p)
    l2, g2 = ops.loss_fwd_bwd("weighted_mse", t, p, global_batch=8)
    assert abs(l1.item() - 2 * l2.item()) < 1e-12
    np.testing.assert_allclose(g1.cpu().numpy(), 2 * g2.cpu().numpy(), rtol=1e-6)


# ---------------------------------------------------------------- scoring
def test_pck_golden(ops, golden_dir):
    g = _load(golden_dir, "score_golden.npz")
    for key, thr in (("pck005", 0.05), ("pck002", 0.02)):
        c, v = ops.pck_counts(g["xs_pred"], g["ys_pred"], g["xs_gt"], g["ys_gt"], g["vs"], g["bbox"][:, 2:4], thr)
        oc, ov = orc.pck_counts(g["xs_pred"], g["ys_pred"], g["xs_gt"], g["ys_gt"], g["vs"], g["bbox"][:, 2:4], thr)
        np.testing.assert_array_equal(c, oc)
        np.testing.assert_array_equal(v, ov)
        np.testing.assert_array_equal(c / v, g[key])


def test_oks_vs_oracle(ops):
    rng = np.random.default_rng(21)
    N = 300
    xg = rng.uniform(0, 400, (N, 17)); yg = rng.uniform(0, 400, (N, 17))
    xp = xg + rng.normal(0, 8, (N, 17)); yp = yg + rng.normal(0, 8, (N, 17))
    vs = rng.integers(0, 3, (N, 17)); vs[:5] = 0
    bb = rng.uniform(10, 300, (N, 4)); area = bb[:, 2] * bb[:, 3] * 0.6
    out = ops.oks_similarity(xp, yp, xg, yg, vs, area, bb).cpu().numpy()
    ref = orc.oks_similarity(xp, yp, xg, yg, vs, area, bb)
    np.testing.assert_allclose(out, ref, rtol=1e-12, atol=1e-15)       # fp64; exp() may differ in the last ulp
