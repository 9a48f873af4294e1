"""Drop-in for the reference's loss.py: same names, same call signature `fn(y_true, y_pred)`, same
returned tensor shapes ((B,H,W) for the MSE family -- "It should NOT return a scalar", loss.py:8-9 --
and (B,) for IOU).  Each function carries `hgb_kind`, which `model.compile(loss=fn)` uses to select
the fused one-pass loss+gradient CUDA kernel (hgb_loss_fwd_bwd); there is no autograd here.
"""
from __future__ import annotations

import numpy as np

from . import _lib, ops


def _map(kind, y_true, y_pred):
    torch = _lib.require_cuda()
    out = ops.loss_map(kind, y_true, y_pred)
    if isinstance(y_pred, torch.Tensor):
        return out
    return out.cpu().numpy()


def weighted_mse(y_true, y_pred):
    """loss.py:2-21: mean over joints of squared error, weighted 82x where the target is > 0."""
    return _map("weighted_mse", y_true, y_pred)


def mean_squared_error(y_true, y_pred):
    """tf.keras.losses.mean_squared_error (trainer.py:231-233)."""
    return _map("mse", y_true, y_pred)


def IOU(y_true, y_pred):
    """loss.py:23-28: 1 - mean_k soft IoU, one value per sample."""
    return _map("iou", y_true, y_pred)


def weighed_keypoint_mse(y_true, y_pred):
    """loss.py:30-36: joints whose target map is all zero do not contribute."""
    return _map("weighted_keypoint_mse", y_true, y_pred)


weighted_mse.hgb_kind = _lib.LOSS_KINDS["weighted_mse"]
mean_squared_error.hgb_kind = _lib.LOSS_KINDS["mse"]
IOU.hgb_kind = _lib.LOSS_KINDS["iou"]
weighed_keypoint_mse.hgb_kind = _lib.LOSS_KINDS["weighted_keypoint_mse"]

_BY_NAME = {
    "weighted_mse": weighted_mse, "weight_mean_squared_error": weighted_mse,
    "mse": mean_squared_error, "mean_squared_error": mean_squared_error,
    "iou": IOU, "weighted_keypoint_mse": weighed_keypoint_mse,
}


def kind_of(loss):
    """Map what was passed to compile(loss=...) to a kernel id; None stays None (predict-only compile)."""
    if loss is None:
        return None
    if isinstance(loss, str):
        fn = _BY_NAME.get(loss.lower())
        if fn is None:
            raise ValueError(f"unknown loss {loss!r}")
        return fn.hgb_kind
    kind = getattr(loss, "hgb_kind", None)
    if kind is None:
        raise TypeError("loss must be one of hgb200.loss.{weighted_mse, mean_squared_error, IOU, weighed_keypoint_mse}: "
                        "the gradient is a hand-written kernel, arbitrary Python callables cannot be differentiated")
    return int(kind)
