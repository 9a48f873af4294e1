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


def crop_and_pad_params(image_height, image_width, square_bbox):
    """The integer bookkeeping of crop_and_pad (data_utils.py:60-96) without touching pixels: returns
    (x0, y0, crop_w, crop_h) such that crop pixel (cy, cx) is source pixel (cy + y0, cx + x0), zero outside the source.
    Raises ValueError where tf.image.crop_to_bounding_box would reject the window."""
    x, y, w, h = square_bbox
    xmin, ymin, xmax, ymax = x, y, x + w, y + h
    offset_width = offset_height = 0
    target_width, target_height = int(image_width), int(image_height)
    if xmin < 0:
        offset_width = int(abs(x))
        target_width += offset_width
    if ymin < 0:
        offset_height = int(abs(y))
        target_height += offset_height
    if xmax > image_width:
        target_width += int(xmax - image_width) + 1
    if ymax > image_height:
        target_height += int(ymax - image_height) + 1
    crop_y, crop_x, crop_h, crop_w = int(max(ymin, 0)), int(max(xmin, 0)), int(h), int(w)
    if crop_w <= 0 or crop_h <= 0:
        raise ValueError("target_width and target_height must be > 0")
    if target_width < crop_w + crop_x:
        raise ValueError("width must be >= target + offset.")
    if target_height < crop_h + crop_y:
        raise ValueError("height must be >= target + offset.")
    return crop_x - offset_width, crop_y - offset_height, crop_w, crop_h


def crop_and_resize(image, square_bboxes, out_height=256, out_width=256):
    """demo.py:44-50 for every box of one frame in ONE launch: uint8 -> float32 conversion, crop_and_pad and
    tf.image.resize(bilinear) fused.  image: (h,w,3) uint8 / float32 host array or CUDA tensor -> (N,out_h,out_w,3) f32 CUDA."""
    h, w = int(image.shape[0]), int(image.shape[1])
    crops = [crop_and_pad_params(h, w, b) for b in square_bboxes]
    return ops.crop_resize([image], crops if crops else np.zeros((0, 4), np.int32), out_height, out_width,
                           source_index=[0] * len(crops))


def crop_and_pad(image, square_bbox):
    """data_utils.py:48-98: the padded crop itself, (int(h), int(w), 3) float32 on device (a resize to its own size is
    the identity: every lerp weight is exactly 0)."""
    x0, y0, cw, ch = crop_and_pad_params(int(image.shape[0]), int(image.shape[1]), square_bbox)
    return ops.crop_resize([image], [(x0, y0, cw, ch)], ch, cw, source_index=[0])[0]


def resize_images(images, out_height=256, out_width=256):
    """tf.image.resize(image, (h, w)) of dataset_builder.py:99,133 for a list of differently sized decoded images."""
    return ops.crop_resize(list(images), None, out_height, out_width)
