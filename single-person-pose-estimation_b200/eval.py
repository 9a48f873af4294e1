"""Drop-in for the reference's eval.py: predict_ds / eval_PCK / eval_OKS with the same signatures and
the same prediction-dict / JSON schema (eval.py:131-139), the arithmetic on the GPU:
  predict_ds  -> model forward, batched heat-map decode on device (only (B,K,3) floats come back)
  eval_PCK    -> hgb_pck_reduce (2*K integer counters)
  eval_OKS    -> COCO keypoint AP/AR: every (detection, ground truth) OKS of the evaluation in one hgb_oks_similarity
                 launch, matching / precision-recall accumulation restated in cocoeval.py (no pycocotools needed).
"""
from __future__ import annotations

import json

import numpy as np

from . import _lib, ops
from .utilities import data_utils


def _np(v):
    if hasattr(v, "detach"):
        v = v.detach().cpu()
    return np.asarray(v.numpy() if hasattr(v, "numpy") else v)


def _load_predictions(path):
    with open(path) as f:
        return json.load(f)


def _save_predictions(predictions, path):
    with open(path, "w") as f:
        json.dump(predictions, f)


def _undo_bbox(x, y, width, height, normalized_xs, normalized_ys):
    """eval.py:153-158."""
    return normalized_xs * width + x, normalized_ys * height + y


def predict_ds(compiled_model, ds, ds_length, batch_size, heatmaps_to_keypoints_func, save_path="result.json",
               conf_threshold=1e-6):
    """eval.py:99-146.  `heatmaps_to_keypoints_func` may be this package's heatmaps_to_keypoints_v1/_v2 (decoded in one
    batched kernel launch on the last stack's device heat maps) or any callable with the reference signature."""
    torch = _lib.require_cuda()
    version = {data_utils.heatmaps_to_keypoints_v1: 1, data_utils.heatmaps_to_keypoints_v2: 2}.get(heatmaps_to_keypoints_func)
    num_iters = int(np.ceil(ds_length / batch_size))
    it = iter(ds)
    predictions = []
    for _ in range(num_iters):
        images_batch, meta = next(it)
        if version is not None and hasattr(compiled_model, "forward_device"):
            x = compiled_model._to_device_images(_np(images_batch) if not isinstance(images_batch, torch.Tensor) else images_batch)
            n = x.shape[0]
            if n < batch_size:
                x = torch.cat([x, x.new_zeros((batch_size - n,) + tuple(x.shape[1:]))], 0)
            last = compiled_model.forward_device(x, training=False)[-1][:n]
            _idx, kp = data_utils.decode_batch(last, conf_threshold, version)
            kpts_all = kp.cpu().numpy()
            hh, ww = last.shape[1], last.shape[2]
        else:
            heat = compiled_model.predict(images_batch)[-1]
            kpts_all = np.stack([heatmaps_to_keypoints_func(hms, conf_threshold=conf_threshold) for hms in heat])
            hh, ww = heat.shape[1], heat.shape[2]
        for j, kpts in enumerate(kpts_all):
            xs_pred = kpts[:, 0] / ww
            ys_pred = kpts[:, 1] / hh
            vs = _np(meta["keypoints/vis"][j])
            bbox_w, bbox_h = int(meta["bbox_w"][j]), int(meta["bbox_h"][j])
            bbox_x, bbox_y = float(meta["bbox_x"][j]), float(meta["bbox_y"][j])
            xs_gt = _np(meta["keypoints/x"][j]) / bbox_w
            ys_gt = _np(meta["keypoints/y"][j]) / bbox_h
            axp, ayp = _undo_bbox(bbox_x, bbox_y, bbox_w, bbox_h, xs_pred, ys_pred)
            axg, ayg = _undo_bbox(bbox_x, bbox_y, bbox_w, bbox_h, xs_gt, ys_gt)
            predictions.append({
                "xs/pred": axp.astype(float).tolist(), "ys/pred": ayp.astype(float).tolist(),
                "xs/gt": axg.astype(float).tolist(), "ys/gt": ayg.astype(float).tolist(),
                "vs": vs.astype(int).tolist(), "confs": kpts[:, 2].astype(float).tolist(),
                "image_id": int(meta["image_id"][j]), "ann_id": int(meta["ann_id"][j]),
                "original_bbox": _np(meta["original_bbox"][j]).astype(float).tolist(),
            })
    _save_predictions(predictions, save_path)
    return predictions


def pck_counts(predictions, pck_threshold=0.05):
    """(correct[K], visible[K]) integer counters of eval.py:62-88, reduced on the GPU."""
    if isinstance(predictions, str):
        predictions = _load_predictions(predictions)
    g = lambda k: np.array([p[k] for p in predictions], dtype=np.float64)  # noqa: E731
    bbox = g("original_bbox")
    return ops.pck_counts(g("xs/pred"), g("ys/pred"), g("xs/gt"), g("ys/gt"),
                          np.array([p["vs"] for p in predictions]), bbox[:, 2:4], pck_threshold)


def eval_PCK(predictions, keypoint_labels, pck_threshold=0.05):
    """eval.py:53-96: fraction of visible joints within pck_threshold * bbox diagonal, per label."""
    correct, visible = pck_counts(predictions, pck_threshold)
    stats = []
    for i, label in enumerate(keypoint_labels):
        percent = int(correct[i]) / int(visible[i])      # ZeroDivisionError for a never-visible joint, as the reference
        stats.append(percent)
        print(f"{label}: {percent:.2f}%")
    return stats


def oks_per_prediction(predictions, areas, bboxes_xywh):
    """COCO OKS of every prediction against its own annotation -> (N,) float64 numpy."""
    if isinstance(predictions, str):
        predictions = _load_predictions(predictions)
    g = lambda k: np.array([p[k] for p in predictions], dtype=np.float64)  # noqa: E731
    out = ops.oks_similarity(g("xs/pred"), g("ys/pred"), g("xs/gt"), g("ys/gt"), np.array([p["vs"] for p in predictions]),
                             np.asarray(areas, np.float64), np.asarray(bboxes_xywh, np.float64))
    return out.cpu().numpy()


def _create_oks_obj(ann_id, image_id, pred_kpts, score):
    return {"image_id": image_id, "ann_id": ann_id, "category_id": 1, "keypoints": pred_kpts, "score": score}


def eval_OKS(predictions, gt_path, backend="hgb", oks_fn=None):
    """eval.py:9-51.  Same call, same printed summary, same `stats[10]` (AP, AP50, AP75, APm, APl, AR, AR50, AR75, ARm, ARl).
    The reference delegates to pycocotools; here the protocol is restated in `cocoeval.py` with every OKS value of the
    evaluation computed by one CUDA launch.  `backend="pycocotools"` runs the third-party package instead (cross-check,
    needs it installed); `oks_fn` replaces the device OKS (tests only)."""
    if backend == "pycocotools":
        from pycocotools.coco import COCO
        from pycocotools.cocoeval import COCOeval
        make_eval = COCOeval
    elif backend == "hgb":
        from .cocoeval import COCO, COCOeval
        make_eval = lambda gt, dt, kind: COCOeval(gt, dt, kind, oks_fn=oks_fn)  # noqa: E731
    else:
        raise ValueError(f"unknown eval_OKS backend {backend!r}")
    if isinstance(predictions, str):
        predictions = _load_predictions(predictions)
    results, image_ids = [], []
    for p in predictions:
        kp = []
        for x, y in zip(p["xs/pred"], p["ys/pred"]):
            kp += [int(x), int(y), 1]                    # coordinates truncated, visibility always 1 (eval.py:25-27)
        results.append(_create_oks_obj(p["ann_id"], p["image_id"], kp, float(np.mean(p["confs"]))))
        image_ids.append(p["image_id"])
    gt = COCO(gt_path)
    ev = make_eval(gt, gt.loadRes(results), "keypoints")
    ev.params.imgIds = image_ids
    ev.params.catIds = [1]
    ev.evaluate()
    ev.accumulate()
    print("\nSummary: ")
    ev.summarize()
    return ev.stats
