"""CPU restatement of the reference's heatmap path -- TEST INFRASTRUCTURE ONLY.

This module is the oracle for target rendering, the losses, heatmap decode and
PCK/OKS scoring.  It is imported only by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  The product package never
imports it; the product fails loudly when the CUDA library is missing.

Parity pin: every function below is checked in tests/test_oracle_golden.py against
fixtures under tests/golden/ that were produced by importing the reference's own
numpy functions from /root/reference (tests/golden/make_golden.py holds the recipe).

Citations are relative to the reference tree.
"""
from __future__ import annotations

import numpy as np

# --------------------------------------------------------------------------- #
# Target rendering  (dataset_builder.py:220-235, utilities/data_utils.py:187-211)
# --------------------------------------------------------------------------- #

def gaussian_patch(sigma: int = 1) -> np.ndarray:
    """7x7 un-normalised Gaussian evaluated in float64 (utilities/data_utils.py:197-202)."""
    size = 6 * sigma + 1
    ax = np.arange(0, size, 1, float)
    c = size // 2
    return np.exp(-((ax[None, :] - c) ** 2 + (ax[:, None] - c) ** 2) / (2 * sigma ** 2))


def render_targets(kps_x, kps_y, kps_v, height: int, width: int) -> np.ndarray:
    """Batch restatement of DatasetBuilder.np_gen_heatmaps (dataset_builder.py:220-235).

    kps_x, kps_y: (B, K) float32 in heatmap pixels; kps_v: (B, K) integer visibility.
    Returns (B, H, W, K) float32.  A joint is drawn iff 0 < int(x) < W, 0 < int(y) < H
    and v > 0 (dataset_builder.py:231); the 7x7 patch is ASSIGNED, clipped to the map
    (utilities/data_utils.py:204-210).  The divide-by-max at :234 is a no-op because the
    centre (value 1.0) is always inside the map.
    """
    kps_x = np.asarray(kps_x, dtype=np.float32)
    kps_y = np.asarray(kps_y, dtype=np.float32)
    kps_v = np.asarray(kps_v)
    B, K = kps_x.shape
    g = gaussian_patch(1)
    out = np.zeros((B, height, width, K), dtype=np.float32)
    for b in range(B):
        for k in range(K):
            x = int(kps_x[b, k])
            y = int(kps_y[b, k])
            if not (0 < x < width and 0 < y < height and kps_v[b, k] > 0):
                continue
            ul = (x - 3, y - 3)
            br = (x + 4, y + 4)
            gx0, gx1 = max(0, -ul[0]), min(br[0], width) - ul[0]
            gy0, gy1 = max(0, -ul[1]), min(br[1], height) - ul[1]
            ix0, ix1 = max(0, ul[0]), min(br[0], width)
            iy0, iy1 = max(0, ul[1]), min(br[1], height)
            out[b, iy0:iy1, ix0:ix1, k] = g[gy0:gy1, gx0:gx1]
    return out


def scale_keypoints(kps, extent, label_extent) -> np.ndarray:
    """dataset_builder.py:107-111 -- two separate float32 ops: divide, then multiply."""
    kps = np.asarray(kps, dtype=np.float32)
    q = (kps / np.float32(extent)).astype(np.float32)
    return (q * np.float32(label_extent)).astype(np.float32)


# --------------------------------------------------------------------------- #
# Losses  (loss.py:2-36, trainer.py:224-245; Keras reduction: SURVEY appendix)
# --------------------------------------------------------------------------- #

def weighted_mse_map(y_true, y_pred) -> np.ndarray:
    """loss.py:2-21 -> (B,H,W): mean over the joint axis of 82/1-weighted squared error."""
    t = np.asarray(y_true, dtype=np.float32)
    p = np.asarray(y_pred, dtype=np.float32)
    w = (t > 0).astype(np.float32) * np.float32(81) + np.float32(1)
    return np.mean(np.square(t - p) * w, axis=-1, dtype=np.float64).astype(np.float32)


def mse_map(y_true, y_pred) -> np.ndarray:
    """tf.keras.losses.mean_squared_error (trainer.py:231-233) -> (B,H,W)."""
    t = np.asarray(y_true, dtype=np.float32)
    p = np.asarray(y_pred, dtype=np.float32)
    return np.mean(np.square(p - t), axis=-1, dtype=np.float64).astype(np.float32)


def keypoint_mse_map(y_true, y_pred) -> np.ndarray:
    """loss.py:30-36 -> (B,H,W): joints whose target map sums to 0 are masked out."""
    t = np.asarray(y_true, dtype=np.float32)
    p = np.asarray(y_pred, dtype=np.float32)
    s = t.sum(axis=(1, 2), keepdims=True)
    kw = 1.0 - (s == 0).astype(np.float32)
    return np.mean(np.square(t - p) * kw, axis=-1, dtype=np.float64).astype(np.float32)


def iou_vec(y_true, y_pred, eps: float = 1e-7) -> np.ndarray:
    """loss.py:23-28 -> (B,): 1 - mean_k soft-IoU."""
    t = np.asarray(y_true, dtype=np.float64)
    p = np.asarray(y_pred, dtype=np.float64)
    inter = (t * p).sum(axis=(1, 2))
    union = (t * t).sum(axis=(1, 2)) + (p * p).sum(axis=(1, 2)) - inter
    iou = (inter + eps) / (union + eps)
    return (1.0 - iou.mean(axis=-1)).astype(np.float32)


LOSS_KINDS = ("weighted_mse", "mse", "iou", "weighted_keypoint_mse")


def loss_and_grad(kind: str, y_true, y_pred):
    """Keras reduction of one output: scalar = mean over everything the loss fn returns;
    also d(scalar)/d(y_pred) in float64 -- the gate for the one-pass CUDA loss kernels."""
    t = np.asarray(y_true, dtype=np.float64)
    p = np.asarray(y_pred, dtype=np.float64)
    B, H, W, K = t.shape
    n = float(B * H * W * K)
    if kind == "weighted_mse":
        w = (t > 0) * 81.0 + 1.0
        return float((w * (t - p) ** 2).sum() / n), 2.0 * w * (p - t) / n
    if kind == "mse":
        return float(((t - p) ** 2).sum() / n), 2.0 * (p - t) / n
    if kind == "weighted_keypoint_mse":
        s32 = np.asarray(y_true, dtype=np.float32).sum(axis=(1, 2), keepdims=True)
        kw = 1.0 - (s32 == 0).astype(np.float64)
        return float((kw * (t - p) ** 2).sum() / n), 2.0 * kw * (p - t) / n
    if kind == "iou":
        eps = 1e-7
        inter = (t * p).sum(axis=(1, 2), keepdims=True)
        tt = (t * t).sum(axis=(1, 2), keepdims=True)
        pp = (p * p).sum(axis=(1, 2), keepdims=True)
        union = tt + pp - inter
        iou = (inter + eps) / (union + eps)
        loss = float((1.0 - iou.mean(axis=-1)).mean())
        # d iou / d p = (t*union - (inter+eps)*(2p - t)) / union^2 ; loss = mean_b(1 - mean_k iou)
        diou = (t * (union + eps) - (inter + eps) * (2.0 * p - t)) / (union + eps) ** 2
        return loss, -diou / float(B * K)
    raise ValueError(kind)


# --------------------------------------------------------------------------- #
# Decode  (utilities/data_utils.py:100-183)
# --------------------------------------------------------------------------- #

def _first_argmax(a: np.ndarray) -> int:
    """np.argmax semantics: first maximal element in C order, a NaN counts as the maximum."""
    return int(np.argmax(a))


def decode_one(heatmap: np.ndarray, conf_threshold: float, version: int):
    """One (H,W) map -> (index, x, y, patch_index, conf, out_x, out_y, out_c).

    v1: utilities/data_utils.py:118-131.  v2 adds :160-182: the clipped 3x3 window, element
    (1,1) of THAT window forced to 0, argmax flattened with the window's true width but
    un-flattened with 3 (:168-169).  The window is not written back here; the caller's-array
    mutation is a host-shim concern.
    """
    H, W = heatmap.shape
    assert H == W, "reference uses index // height (data_utils.py:122): square maps only"
    index = _first_argmax(heatmap)
    x = index % W
    y = index // H
    conf = heatmap[y, x]
    pidx = 0
    if version == 2:
        x1, x2 = max(x - 1, 0), min(x + 2, W)
        y1, y2 = max(y - 1, 0), min(y + 2, H)
        patch = np.array(heatmap[y1:y2, x1:x2], copy=True)
        patch[1][1] = 0
        pidx = _first_argmax(patch)
    px, py = pidx % 3, pidx // 3
    if conf > conf_threshold:
        ox, oy, oc = np.float32(x + px / 4), np.float32(y + py / 4), np.float32(conf)
    else:
        ox = oy = oc = np.float32(0)
    return index, x, y, pidx, np.float32(conf), ox, oy, oc


def decode_batch(heatmaps: np.ndarray, conf_threshold: float = 1e-6, version: int = 2):
    """(B,H,W,K) -> (idx int32 (B,K,4) = [index, x, y, patch_index], kpts float32 (B,K,3))."""
    hm = np.asarray(heatmaps)
    B, H, W, K = hm.shape
    idx = np.zeros((B, K, 4), dtype=np.int32)
    out = np.zeros((B, K, 3), dtype=np.float32)
    for b in range(B):
        for k in range(K):
            i, x, y, pi, _c, ox, oy, oc = decode_one(hm[b, :, :, k].astype(np.float32), conf_threshold, version)
            idx[b, k] = (i, x, y, pi)
            out[b, k] = (ox, oy, oc)
    return idx, out


# --------------------------------------------------------------------------- #
# Scoring  (eval.py:53-96 PCK in-repo; OKS = public COCO keypoint similarity that
# pycocotools.COCOeval.computeOks applies -- third party, unpinned: "parity unpinned")
# --------------------------------------------------------------------------- #

def pck_counts(xs_pred, ys_pred, xs_gt, ys_gt, vs, bbox_wh, pck_threshold: float = 0.05):
    """eval.py:62-88 as integer counters: (correct[K], visible[K]); float64 like the reference."""
    xs_pred = np.asarray(xs_pred, dtype=np.float64)
    ys_pred = np.asarray(ys_pred, dtype=np.float64)
    xs_gt = np.asarray(xs_gt, dtype=np.float64)
    ys_gt = np.asarray(ys_gt, dtype=np.float64)
    vs = np.asarray(vs)
    bbox_wh = np.asarray(bbox_wh, dtype=np.float64)
    N, K = xs_pred.shape
    correct = np.zeros(K, dtype=np.int64)
    visible = np.zeros(K, dtype=np.int64)
    for n in range(N):
        diameter = np.sqrt(bbox_wh[n, 0] ** 2 + bbox_wh[n, 1] ** 2)
        threshold = pck_threshold * diameter
        for k in range(K):
            if vs[n, k] > 0:
                dist = np.sqrt((xs_gt[n, k] - xs_pred[n, k]) ** 2 + (ys_gt[n, k] - ys_pred[n, k]) ** 2)
                visible[k] += 1
                if dist <= threshold:
                    correct[k] += 1
    return correct, visible


COCO_SIGMAS = np.array([.26, .25, .25, .35, .35, .79, .79, .72, .72, .62, .62,
                        1.07, 1.07, .87, .87, .89, .89]) / 10.0


def oks_similarity(xs_pred, ys_pred, xs_gt, ys_gt, vs, area, bbox_xywh):
    """Public COCO object-keypoint-similarity between each prediction and its own GT
    annotation (the arithmetic eval.py:39-49 delegates to pycocotools).  (N,) float64.
    Zero visible joints -> distance to the doubled bbox, averaged over all joints."""
    xs_pred = np.asarray(xs_pred, dtype=np.float64)
    ys_pred = np.asarray(ys_pred, dtype=np.float64)
    xs_gt = np.asarray(xs_gt, dtype=np.float64)
    ys_gt = np.asarray(ys_gt, dtype=np.float64)
    vs = np.asarray(vs)
    area = np.asarray(area, dtype=np.float64)
    bb = np.asarray(bbox_xywh, dtype=np.float64)
    N, K = xs_pred.shape
    vars_ = (COCO_SIGMAS[:K] * 2) ** 2
    out = np.zeros(N, dtype=np.float64)
    for n in range(N):
        k1 = int(np.count_nonzero(vs[n] > 0))
        if k1 > 0:
            dx = xs_pred[n] - xs_gt[n]
            dy = ys_pred[n] - ys_gt[n]
        else:
            x0, x1 = bb[n, 0] - bb[n, 2], bb[n, 0] + bb[n, 2] * 2
            y0, y1 = bb[n, 1] - bb[n, 3], bb[n, 1] + bb[n, 3] * 2
            z = np.zeros(K)
            dx = np.maximum(z, x0 - xs_pred[n]) + np.maximum(z, xs_pred[n] - x1)
            dy = np.maximum(z, y0 - ys_pred[n]) + np.maximum(z, ys_pred[n] - y1)
        e = (dx ** 2 + dy ** 2) / vars_ / (area[n] + np.spacing(1)) / 2
        if k1 > 0:
            e = e[vs[n] > 0]
        out[n] = np.sum(np.exp(-e)) / e.shape[0]
    return out
