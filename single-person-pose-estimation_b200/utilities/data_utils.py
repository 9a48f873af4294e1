"""Drop-in for the hot-path functions of the reference's utilities/data_utils.py.

heatmaps_to_keypoints_v1/_v2 keep the reference signature (one (H,W,K) array in, (K,3) float32 out)
but run the CUDA decode kernel (hgb_decode); `decode_batch` is the batched device entry the fast
predict loop uses.  v2 also reproduces the reference's side effect: element (1,1) of the clipped 3x3
window around each peak is set to 0 in the CALLER's array (data_utils.py:165-166).
"""
from __future__ import annotations

import numpy as np

from .. import ops


def decode_batch(heatmaps, conf_threshold=1e-6, version=2):
    """(B,H,W,K) device or host array -> (idx int32 (B,K,4) = [argmax, x, y, patch argmax], kpts f32 (B,K,3)), on device."""
    return ops.decode_batch(heatmaps, conf_threshold, version)


def _decode_one(heatmaps, conf_threshold, version):
    hm = np.asarray(heatmaps) if not hasattr(heatmaps, "is_cuda") else heatmaps
    if hm.ndim != 3:
        raise ValueError("heatmaps must be (height, width, num_kps)")
    idx, kp = ops.decode_batch(hm[None], conf_threshold, version)
    return idx[0].cpu().numpy(), kp[0].cpu().numpy()


def heatmaps_to_keypoints_v1(heatmaps, conf_threshold=1e-6):
    """Per-joint argmax -> (x, y, confidence), zeros below the threshold (data_utils.py:100-132)."""
    return _decode_one(heatmaps, conf_threshold, 1)[1]


def heatmaps_to_keypoints_v2(heatmaps, conf_threshold=1e-6):
    """v1 plus the quarter-pixel offset towards the second-highest value of the 3x3 window
    (data_utils.py:135-183), including the in-place zeroing of the window's (1,1) element."""
    idx, kp = _decode_one(heatmaps, conf_threshold, 2)
    if isinstance(heatmaps, np.ndarray) and heatmaps.flags.writeable:
        ys = np.maximum(idx[:, 2] - 1, 0) + 1
        xs = np.maximum(idx[:, 1] - 1, 0) + 1
        heatmaps[ys, xs, np.arange(heatmaps.shape[2])] = 0
    return kp


def transform_bbox_square(bbox, scale=1):
    """Centre-preserving square box with side = longer side * scale (data_utils.py:16-45)."""
    x, y, w, h = bbox
    cx, cy = x + w / 2, y + h / 2
    side_w = side_h = w if w >= h else h
    side_w *= scale
    side_h *= scale
    return cx - side_w / 2, cy - side_h / 2, side_w, side_h
