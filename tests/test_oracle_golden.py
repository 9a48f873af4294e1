"""Pin the oracle (oracle/heatmap_oracle.py) to fixtures produced by the reference's own
numpy functions (tests/golden/make_golden.py).  CPU only."""
import os

import numpy as np

from oracle import heatmap_oracle as orc


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


def test_decode_matches_reference(golden_dir):
    g = _load(golden_dir, "decode_golden.npz")
    hm = g["heatmaps"]
    for ti, thr in enumerate(g["thresholds"]):
        for ver in (1, 2):
            _idx, out = orc.decode_batch(hm, float(thr), ver)
            ref = g[f"v{ver}_thr{ti}"]
            assert out.dtype == np.float32
            np.testing.assert_array_equal(out.view(np.uint32), ref.view(np.uint32))


def test_decode_offsets_are_quarter_steps(golden_dir):
    g = _load(golden_dir, "decode_golden.npz")
    v2 = g["v2_thr0"]
    frac = v2[..., :2] - np.floor(v2[..., :2])
    assert set(np.unique(frac)).issubset({0.0, 0.25, 0.5})


def test_render_matches_reference(golden_dir):
    g = _load(golden_dir, "render_golden.npz")
    out = orc.render_targets(g["kps_x"], g["kps_y"], g["kps_v"], 64, 64)
    np.testing.assert_array_equal(out.view(np.uint32), g["targets"].view(np.uint32))
    out128 = orc.render_targets(g["kps_x128"], g["kps_y128"], g["kps_v"][:3], 128, 128)
    ref128 = np.zeros((3, 128, 128, 17), np.float32)
    ref128[tuple(g["t128_idx"])] = g["t128_val"]
    np.testing.assert_array_equal(out128.view(np.uint32), ref128.view(np.uint32))
    np.testing.assert_array_equal(orc.gaussian_patch(1).astype(np.float32), g["gaussian7"])


def test_render_edge_rules(golden_dir):
    g = _load(golden_dir, "render_golden.npz")
    t = g["targets"][0]
    assert np.count_nonzero(t[:, :, 0]) == 49 and t[20, 10, 0] == 1.0      # (10.7,20.2)
    assert np.count_nonzero(t[:, :, 1]) == 25                              # (1.9,62.5) clipped
    assert np.count_nonzero(t[:, :, 2]) == 0                               # x=0.5 -> int 0 rejected
    assert np.count_nonzero(t[:, :, 3]) == 16 and t[63, 63, 3] == 1.0      # (63.9,63.9)
    assert len(np.unique(t[:, :, 0])) == 11                                # 10 non-zero values + 0


def test_pck_matches_reference(golden_dir):
    g = _load(golden_dir, "score_golden.npz")
    for key, thr in (("pck005", 0.05), ("pck002", 0.02)):
        c, v = orc.pck_counts(g["xs_pred"], g["ys_pred"], g["xs_gt"], g["ys_gt"], g["vs"], g["bbox"][:, 2:4], thr)
        np.testing.assert_array_equal(c / v, g[key])


def test_losses_consistent():
    rng = np.random.default_rng(0)
    kx = rng.uniform(-4, 68, (3, 17)).astype(np.float32)
    ky = rng.uniform(-4, 68, (3, 17)).astype(np.float32)
    kv = rng.integers(0, 3, (3, 17))
    t = orc.render_targets(kx, ky, kv, 64, 64)
    p = (1 / (1 + np.exp(-rng.standard_normal(t.shape)))).astype(np.float32)
    for kind, fn in (("weighted_mse", orc.weighted_mse_map), ("mse", orc.mse_map),
                     ("weighted_keypoint_mse", orc.keypoint_mse_map), ("iou", orc.iou_vec)):
        loss, grad = orc.loss_and_grad(kind, t, p)
        assert abs(loss - float(np.mean(fn(t, p), dtype=np.float64))) < 1e-6 * max(1.0, abs(loss))
        # finite-difference check of the analytic gradient on a few coordinates
        for _ in range(5):
            i = tuple(rng.integers(0, s) for s in t.shape)
            q = p.astype(np.float64).copy()
            h = 1e-4
            q[i] += h
            lp, _ = orc.loss_and_grad(kind, t, q)
            q[i] -= 2 * h
            lm, _ = orc.loss_and_grad(kind, t, q)
            fd = (lp - lm) / (2 * h)
            assert abs(fd - grad[i]) <= 1e-6 + 1e-3 * abs(fd)
    assert orc.weighted_mse_map(t, p).shape == (3, 64, 64)
    assert orc.iou_vec(t, p).shape == (3,)


def test_keypoint_scaling_two_ops():
    x = np.array([100.0, 33.3, 250.7], np.float32)
    out = orc.scale_keypoints(x, 301, 64)
    exp = (x / np.float32(301)) * np.float32(64)
    np.testing.assert_array_equal(out, exp.astype(np.float32))
