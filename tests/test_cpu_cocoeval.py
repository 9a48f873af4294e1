"""COCO keypoint AP/AR protocol (cocoeval.py, replaces the pycocotools calls of eval.py:39-49): host logic on the
CPU with the numpy OKS oracle injected.  Expected values come from an independent, deliberately naive restatement of
the published protocol below (single ground truth per image) and from hand-derived cases."""
import json

import numpy as np
import pytest

from oracle import heatmap_oracle as horc


def _person(ann_id, image_id, xy, vis, iscrowd=0, area=None, bbox=None):
    xy = np.asarray(xy, float)
    kp = []
    for (x, y), v in zip(xy, vis):
        kp += [float(x), float(y), int(v)]
    x0, y0, x1, y1 = xy[:, 0].min(), xy[:, 1].min(), xy[:, 0].max(), xy[:, 1].max()
    bbox = bbox or [x0, y0, x1 - x0, y1 - y0]
    return {"id": ann_id, "image_id": image_id, "category_id": 1, "keypoints": kp, "num_keypoints": int(np.sum(np.asarray(vis) > 0)),
            "iscrowd": iscrowd, "area": float(area if area is not None else bbox[2] * bbox[3]), "bbox": [float(b) for b in bbox]}


def _dataset(anns):
    return {"images": [{"id": i} for i in sorted({a["image_id"] for a in anns})], "annotations": anns,
            "categories": [{"id": 1, "name": "person"}]}


def _skeleton(rng, scale=200.0, offset=50.0):
    return rng.random((17, 2)) * scale + offset


def _prediction(ann, xy, conf):
    return {"xs/pred": [float(v) for v in xy[:, 0]], "ys/pred": [float(v) for v in xy[:, 1]], "confs": [float(conf)] * 17,
            "image_id": ann["image_id"], "ann_id": ann["id"]}


def _naive_ap(oks, scores, thr, gt_in_range=None, dt_in_range=None):
    """101-point interpolated AP / final recall at one threshold for images holding one ground truth and one detection.
    A detection is ignored when it matches an out-of-range ground truth, or matches nothing and is itself out of range."""
    n = len(oks)
    gt_in = np.ones(n, bool) if gt_in_range is None else np.asarray(gt_in_range)
    dt_in = np.ones(n, bool) if dt_in_range is None else np.asarray(dt_in_range)
    matched = np.asarray(oks) >= thr
    ignored = np.where(matched, ~gt_in, ~dt_in)
    order = [i for i in np.argsort(-np.asarray(scores), kind="mergesort") if not ignored[i]]
    tp = matched[order].astype(float)
    n_gt = int(gt_in.sum())
    if len(order) == 0:
        return 0.0, 0.0
    ctp, cfp = np.cumsum(tp), np.cumsum(1 - tp)
    rc, pr = ctp / n_gt, ctp / (ctp + cfp + np.spacing(1))
    for i in range(len(pr) - 1, 0, -1):
        pr[i - 1] = max(pr[i - 1], pr[i])
    ap = 0.0
    for r in np.linspace(0, 1, 101):
        idx = np.searchsorted(rc, r, side="left")
        ap += pr[idx] if idx < len(pr) else 0.0
    return ap / 101, rc[-1]


@pytest.fixture()
def hgb():
    import hgb200
    return hgb200


def test_perfect_predictions_score_one(hgb, tmp_path, capsys):
    rng = np.random.default_rng(0)
    anns = [_person(100 + i, 10 + i, _skeleton(rng), [2] * 17) for i in range(5)]
    gt_path = tmp_path / "gt.json"
    gt_path.write_text(json.dumps(_dataset(anns)))
    preds = [_prediction(a, np.asarray(a["keypoints"]).reshape(17, 3)[:, :2], 0.9) for a in anns]
    stats = hgb.eval.eval_OKS(preds, str(gt_path), oks_fn=horc.oks_similarity)
    out = capsys.readouterr().out
    assert "Summary: " in out and "Average Precision  (AP) @[ IoU=0.50:0.95 | area=   all | maxDets= 20 ] = 1.000" in out
    assert stats.shape == (10,)
    big = [a["area"] > 96 ** 2 for a in anns]
    assert stats[0] == stats[1] == stats[2] == stats[5] == stats[6] == stats[7] == 1.0
    assert stats[4] == (1.0 if any(big) else -1)


def test_matches_naive_protocol_on_random_single_person_images(hgb, tmp_path):
    rng = np.random.default_rng(1)
    n = 40
    anns, preds, scores = [], [], []
    for i in range(n):
        sk = _skeleton(rng)
        vis = rng.integers(0, 3, 17)
        vis[0] = 2
        a = _person(1000 + i, 1 + i, sk, vis, area=rng.uniform(40 ** 2, 200 ** 2))
        anns.append(a)
        noise = rng.normal(0, rng.choice([1.0, 6.0, 25.0]), (17, 2))
        conf = float(rng.uniform(0.05, 0.95))
        preds.append(_prediction(a, sk + noise, conf))
        scores.append(conf)
    gt_path = tmp_path / "gt.json"
    gt_path.write_text(json.dumps(_dataset(anns)))
    # what the protocol sees: truncated coordinates (eval.py:25-26)
    oks = []
    for a, p in zip(anns, preds):
        g = np.asarray(a["keypoints"]).reshape(17, 3)
        xs, ys = np.trunc(p["xs/pred"]), np.trunc(p["ys/pred"])
        oks.append(horc.oks_similarity(xs[None], ys[None], g[None, :, 0], g[None, :, 1], g[None, :, 2], [a["area"]], [a["bbox"]])[0])
    stats = hgb.eval.eval_OKS(preds, str(gt_path), oks_fn=horc.oks_similarity)
    thrs = np.linspace(.5, .95, 10)
    aps, ars = zip(*[_naive_ap(oks, scores, t) for t in thrs])
    assert stats[0] == pytest.approx(np.mean(aps), abs=1e-12)
    assert stats[1] == pytest.approx(aps[0], abs=1e-12)
    assert stats[2] == pytest.approx(aps[5], abs=1e-12)
    assert stats[5] == pytest.approx(np.mean(ars), abs=1e-12)
    assert stats[6] == pytest.approx(ars[0], abs=1e-12)
    assert stats[7] == pytest.approx(ars[5], abs=1e-12)
    assert 0 < stats[0] < 1                                   # the case is not degenerate
    # area ranges: ground truths by their annotated area, detections by the extent of their (truncated) joints
    dt_area = [(np.trunc(p["xs/pred"]).max() - np.trunc(p["xs/pred"]).min()) * (np.trunc(p["ys/pred"]).max() - np.trunc(p["ys/pred"]).min())
               for p in preds]
    for lo, hi, ap_i, ar_i in ((32 ** 2, 96 ** 2, 3, 8), (96 ** 2, 1e10, 4, 9)):
        gin = [lo <= a["area"] <= hi for a in anns]
        din = [lo <= d <= hi for d in dt_area]
        assert any(gin)
        aps, ars = zip(*[_naive_ap(oks, scores, t, gin, din) for t in thrs])
        assert stats[ap_i] == pytest.approx(np.mean(aps), abs=1e-12)
        assert stats[ar_i] == pytest.approx(np.mean(ars), abs=1e-12)


def test_greedy_matching_ignore_rules_and_score_order(hgb):
    from hgb200.cocoeval import COCO, COCOeval
    rng = np.random.default_rng(2)
    a, b, c = _skeleton(rng), _skeleton(rng, offset=400.0), _skeleton(rng, offset=800.0)
    anns = [_person(1, 7, a, [2] * 17, area=150 ** 2), _person(2, 7, b, [2] * 17, area=150 ** 2),
            _person(3, 7, c, [0] * 17, area=150 ** 2)]            # no labelled joints -> ignored
    gt = COCO(_dataset(anns))

    def det(xy, score):
        kp = []
        for x, y in xy:
            kp += [float(x), float(y), 1]
        return {"image_id": 7, "category_id": 1, "keypoints": kp, "score": score}

    # two detections on person 1 (the lower-scored one more accurate), one on person 2, one inside the ignored box, one nowhere
    dets = [det(a + 3.0, 0.9), det(a, 0.8), det(b, 0.7), det(c, 0.6), det(a * 0 + 5000.0, 0.5)]
    ev = COCOeval(gt, gt.loadRes(dets), "keypoints", oks_fn=horc.oks_similarity)
    ev.params.imgIds, ev.params.catIds = [7], [1]
    ev.evaluate()
    e = ev.evalImgs[0]                                             # area range 'all'
    assert e["dtScores"] == [0.9, 0.8, 0.7, 0.6, 0.5]
    assert e["gtIds"] == [1, 2, 3] and list(e["gtIgnore"]) == [0, 0, 1]
    # highest score claims person 1 first even though detection 2 fits better; detection 2 is then unmatched
    assert list(e["dtMatches"][0]) == [1, 0, 2, 3, 0]
    assert list(e["dtIgnore"][0]) == [False, False, False, True, False]
    ev.accumulate()
    ev.summarize()
    # at OKS .5: order tp fp tp (ignored) fp ; 2 regular gts -> recall 1 reached at the third detection with precision 2/3
    pr = ev.eval["precision"][0, :, 0, 0, 0]
    assert pr[0] == pytest.approx(1.0) and pr[50] == pytest.approx(1.0) and pr[51] == pytest.approx(2 / 3) and pr[100] == pytest.approx(2 / 3)
    assert ev.eval["recall"][0, 0, 0, 0] == 1.0
    assert ev.stats[1] == pytest.approx((51 * 1.0 + 50 * 2 / 3) / 101)


def test_crowd_regions_can_absorb_several_detections_and_empty_cases(hgb):
    from hgb200.cocoeval import COCO, COCOeval
    rng = np.random.default_rng(3)
    a = _skeleton(rng)
    anns = [_person(1, 1, a, [2] * 17, iscrowd=1, area=100 ** 2), _person(2, 2, a, [2] * 17, area=100 ** 2)]
    ds = _dataset(anns)
    ds["images"].append({"id": 3})
    gt = COCO(ds)
    kp = []
    for x, y in a:
        kp += [float(x), float(y), 1]
    dets = [{"image_id": 1, "category_id": 1, "keypoints": kp, "score": 0.9},
            {"image_id": 1, "category_id": 1, "keypoints": kp, "score": 0.8},
            {"image_id": 3, "category_id": 1, "keypoints": kp, "score": 0.95}]      # image without ground truth: false positive
    ev = COCOeval(gt, gt.loadRes(dets), "keypoints", oks_fn=horc.oks_similarity)
    ev.evaluate()
    by_img = {e["image_id"]: e for e in ev.evalImgs[:3] if e is not None}
    assert by_img[1]["dtIgnore"][0].tolist() == [True, True]      # both land on the crowd region and are ignored
    assert by_img[2]["dtIds"] == [] and by_img[2]["gtIds"] == [2]
    assert by_img[3]["gtIds"] == [] and by_img[3]["dtMatches"].shape == (10, 1)
    ev.accumulate()
    ev.summarize()
    assert ev.stats[0] == 0.0 and ev.stats[5] == 0.0               # one regular gt, never found; one false positive


def test_loadres_contract_and_param_table(hgb):
    from hgb200.cocoeval import COCO, Params
    gt = COCO(_dataset([_person(1, 1, _skeleton(np.random.default_rng(4)), [2] * 17)]))
    res = gt.loadRes([{"image_id": 1, "category_id": 1, "keypoints": [10, 20, 1, 30, 60, 1] + [10, 20, 1] * 15, "score": 1.0}])
    ann = res.loadAnns(res.getAnnIds(imgIds=[1]))[0]
    assert ann["id"] == 1 and ann["bbox"] == [10, 20, 20, 40] and ann["area"] == 800
    with pytest.raises(AssertionError):
        gt.loadRes([{"image_id": 99, "category_id": 1, "keypoints": [0] * 51, "score": 1.0}])
    p = Params()
    assert len(p.iouThrs) == 10 and p.iouThrs[0] == 0.5 and p.iouThrs[5] == 0.75 and len(p.recThrs) == 101
    assert p.maxDets == [20] and p.areaRngLbl == ["all", "medium", "large"]
    with pytest.raises(ValueError):
        Params("bbox")


def test_default_oks_is_the_cuda_kernel_and_fails_loudly_without_gpu(hgb):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from hgb200.cocoeval import COCO, COCOeval
    a = _skeleton(np.random.default_rng(5))
    gt = COCO(_dataset([_person(1, 1, a, [2] * 17)]))
    kp = []
    for x, y in a:
        kp += [float(x), float(y), 1]
    ev = COCOeval(gt, gt.loadRes([{"image_id": 1, "category_id": 1, "keypoints": kp, "score": 1.0}]), "keypoints")
    with pytest.raises(hgb._lib.HgbError):
        ev.evaluate()
