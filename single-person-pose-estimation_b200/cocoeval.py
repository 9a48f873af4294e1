"""COCO keypoint evaluation (AP / AR over OKS thresholds) without pycocotools.

The reference's eval_OKS (eval.py:9-51) hands its predictions to pycocotools: `COCO(gt_path)`, `loadRes`,
`COCOeval(gt, dt, 'keypoints')`, `evaluate()`, `accumulate()`, `summarize()` -> `stats[10]`.  pycocotools is a
third-party dependency that is not vendored in the reference (and absent from this image), so this module
restates the *published COCO keypoint protocol* those calls implement, with the same class / method /
attribute names, so `eval_OKS` keeps working and the two can be swapped:

  * OKS between every detection and every ground-truth person of the same image (sigma table, doubled-bbox
    rule for annotations without labelled joints) -- this is the arithmetic, and it runs on the GPU: all
    (detection, ground truth) pairs of the whole evaluation are flattened into ONE `hgb_oks_similarity`
    launch (`ops.oks_similarity`); nothing is computed per image on the host.
  * greedy matching per image / area range / OKS threshold, detections in descending score order, ignored
    ground truths (crowd, no labelled joints, outside the area range) matched last;
  * precision at 101 recall thresholds with the monotone envelope, averaged into AP / AP50 / AP75 / APm /
    APl and AR / AR50 / AR75 / ARm / ARl.

The matching and the precision/recall accumulation are host bookkeeping over a few integers per detection
(exactly as in the reference's dependency); they carry no floating-point arithmetic beyond comparisons and
one division per recall point.

`oks_fn` can be injected (tests run the host logic on machines without a GPU by passing the numpy oracle);
the default is the CUDA kernel and raises when CUDA is unavailable -- there is no CPU fallback.
"""
from __future__ import annotations

import copy
import json
from collections import defaultdict

import numpy as np

KPT_OKS_SIGMAS = np.array([.26, .25, .25, .35, .35, .79, .79, .72, .72, .62, .62, 1.07, 1.07, .87, .87, .89, .89]) / 10.0


def _device_oks(xs_pred, ys_pred, xs_gt, ys_gt, vs, area, bbox_xywh):
    from . import ops
    return ops.oks_similarity(xs_pred, ys_pred, xs_gt, ys_gt, vs, area, bbox_xywh).cpu().numpy()


class COCO:
    """The slice of pycocotools.coco.COCO that keypoint evaluation touches: annotation file in, id-indexed
    lookups, and `loadRes` (eval.py:39-40)."""

    def __init__(self, annotation_file=None):
        self.dataset, self.anns, self.imgs, self.cats = {}, {}, {}, {}
        self.imgToAnns, self.catToImgs = defaultdict(list), defaultdict(list)
        if annotation_file is not None:
            if isinstance(annotation_file, dict):
                dataset = annotation_file
            else:
                with open(annotation_file) as f:
                    dataset = json.load(f)
            if not isinstance(dataset, dict):
                raise AssertionError(f"annotation file format {type(dataset)} not supported")
            self.dataset = dataset
            self.createIndex()

    def createIndex(self):
        anns, imgs, cats = {}, {}, {}
        img_to_anns, cat_to_imgs = defaultdict(list), defaultdict(list)
        for ann in self.dataset.get("annotations", []):
            img_to_anns[ann["image_id"]].append(ann)
            anns[ann["id"]] = ann
            cat_to_imgs[ann["category_id"]].append(ann["image_id"])
        for img in self.dataset.get("images", []):
            imgs[img["id"]] = img
        for cat in self.dataset.get("categories", []):
            cats[cat["id"]] = cat
        self.anns, self.imgs, self.cats = anns, imgs, cats
        self.imgToAnns, self.catToImgs = img_to_anns, cat_to_imgs

    @staticmethod
    def _as_list(v):
        return list(v) if isinstance(v, (list, tuple, set, np.ndarray)) else [v]

    def getAnnIds(self, imgIds=(), catIds=(), iscrowd=None):
        img_ids, cat_ids = self._as_list(imgIds), self._as_list(catIds)
        if img_ids:
            anns = [a for i in img_ids if i in self.imgToAnns for a in self.imgToAnns[i]]
        else:
            anns = self.dataset.get("annotations", [])
        if cat_ids:
            cat_set = set(cat_ids)
            anns = [a for a in anns if a["category_id"] in cat_set]
        if iscrowd is not None:
            anns = [a for a in anns if a["iscrowd"] == iscrowd]
        return [a["id"] for a in anns]

    def getImgIds(self):
        return list(self.imgs.keys())

    def getCatIds(self):
        return [c["id"] for c in self.dataset.get("categories", [])]

    def loadAnns(self, ids=()):
        return [self.anns[i] for i in self._as_list(ids)]

    def loadRes(self, resFile):
        """Result list (or json path) -> COCO object.  Keypoint results get `area` / `bbox` from the extent of
        their coordinates and ids 1..n, as the COCO result format prescribes."""
        res = COCO()
        res.dataset["images"] = list(self.dataset.get("images", []))
        if isinstance(resFile, str):
            with open(resFile) as f:
                anns = json.load(f)
        else:
            anns = resFile
        if not isinstance(anns, list):
            raise AssertionError("results is not an array of objects")
        anns = copy.deepcopy(anns)
        missing = set(a["image_id"] for a in anns) - set(self.getImgIds())
        if missing:
            raise AssertionError("Results do not correspond to current coco set")
        if anns and "keypoints" not in anns[0]:
            raise AssertionError("only keypoint results are supported on this path (eval.py:42)")
        res.dataset["categories"] = copy.deepcopy(self.dataset.get("categories", []))
        for n, ann in enumerate(anns):
            s = ann["keypoints"]
            x, y = s[0::3], s[1::3]
            x0, x1, y0, y1 = min(x), max(x), min(y), max(y)
            ann["area"] = (x1 - x0) * (y1 - y0)
            ann["id"] = n + 1
            ann["bbox"] = [x0, y0, x1 - x0, y1 - y0]
        res.dataset["annotations"] = anns
        res.createIndex()
        return res


class Params:
    """Evaluation parameters of the keypoint protocol."""

    def __init__(self, iouType="keypoints"):
        if iouType != "keypoints":
            raise ValueError("only iouType='keypoints' is on this path (eval.py:42)")
        self.iouType = iouType
        self.imgIds, self.catIds = [], []
        self.iouThrs = np.linspace(.5, 0.95, int(np.round((0.95 - .5) / .05)) + 1, endpoint=True)
        self.recThrs = np.linspace(.0, 1.00, int(np.round((1.00 - .0) / .01)) + 1, endpoint=True)
        self.maxDets = [20]
        self.areaRng = [[0 ** 2, 1e5 ** 2], [32 ** 2, 96 ** 2], [96 ** 2, 1e5 ** 2]]
        self.areaRngLbl = ["all", "medium", "large"]
        self.useCats = 1
        self.kpt_oks_sigmas = KPT_OKS_SIGMAS.copy()


class COCOeval:
    def __init__(self, cocoGt=None, cocoDt=None, iouType="keypoints", oks_fn=None):
        self.cocoGt, self.cocoDt = cocoGt, cocoDt
        self.params = Params(iouType)
        self.evalImgs, self.eval, self.ious = [], {}, {}
        self._gts, self._dts = defaultdict(list), defaultdict(list)
        self._paramsEval = None
        self.stats = []
        self._oks_fn = oks_fn or _device_oks
        if cocoGt is not None:
            self.params.imgIds = sorted(cocoGt.getImgIds())
            self.params.catIds = sorted(cocoGt.getCatIds())

    # ------------------------------------------------------------------ per-image preparation
    def _prepare(self):
        p = self.params
        cat_ids = p.catIds if p.useCats else []
        gts = self.cocoGt.loadAnns(self.cocoGt.getAnnIds(imgIds=p.imgIds, catIds=cat_ids))
        dts = self.cocoDt.loadAnns(self.cocoDt.getAnnIds(imgIds=p.imgIds, catIds=cat_ids))
        for gt in gts:
            crowd = bool("iscrowd" in gt and gt["iscrowd"])
            nk = gt["num_keypoints"] if "num_keypoints" in gt else int(np.count_nonzero(np.asarray(gt["keypoints"])[2::3] > 0))
            gt["ignore"] = (nk == 0) or crowd
        self._gts, self._dts = defaultdict(list), defaultdict(list)
        for gt in gts:
            self._gts[gt["image_id"], gt["category_id"]].append(gt)
        for dt in dts:
            self._dts[dt["image_id"], dt["category_id"]].append(dt)
        self.evalImgs, self.eval = [], {}

    def _top_dets(self, img_id, cat_id):
        dts = self._dts[img_id, cat_id]
        order = np.argsort([-d["score"] for d in dts], kind="mergesort")
        return [dts[i] for i in order][: self.params.maxDets[-1]]

    def _compute_all_oks(self, keys):
        """One device launch for every (detection, ground truth) pair of the evaluation."""
        k = len(self.params.kpt_oks_sigmas)
        if k != len(KPT_OKS_SIGMAS) or not np.array_equal(self.params.kpt_oks_sigmas, KPT_OKS_SIGMAS):
            raise ValueError("the OKS kernel carries the 17 COCO person sigmas; custom sigmas are not supported")
        shapes, rows_d, rows_g, vis, area, bbox = {}, [], [], [], [], []
        for key in keys:
            gts, dts = self._gts[key], self._top_dets(*key)
            if not gts or not dts:
                shapes[key] = None
                continue
            shapes[key] = (len(dts), len(gts))
            for dt in dts:
                d = np.asarray(dt["keypoints"], dtype=np.float64)
                for gt in gts:
                    g = np.asarray(gt["keypoints"], dtype=np.float64)
                    rows_d.append(d)
                    rows_g.append(g)
                    vis.append(g[2::3] > 0)
                    area.append(gt["area"])
                    bbox.append(gt["bbox"])
        if rows_d:
            d, g = np.stack(rows_d), np.stack(rows_g)
            flat = np.asarray(self._oks_fn(d[:, 0::3], d[:, 1::3], g[:, 0::3], g[:, 1::3],
                                           np.stack(vis).astype(np.int32), np.asarray(area, np.float64),
                                           np.asarray(bbox, np.float64)), dtype=np.float64)
        else:
            flat = np.zeros(0)
        out, at = {}, 0
        for key in keys:
            if shapes[key] is None:
                out[key] = []
            else:
                n = shapes[key][0] * shapes[key][1]
                out[key] = flat[at:at + n].reshape(shapes[key])
                at += n
        return out

    # ------------------------------------------------------------------ evaluate
    def evaluate(self):
        p = self.params
        print("Running per image evaluation...")
        print(f"Evaluate annotation type *{p.iouType}*")
        p.imgIds = list(np.unique(p.imgIds))
        if p.useCats:
            p.catIds = list(np.unique(p.catIds))
        p.maxDets = sorted(p.maxDets)
        self._prepare()
        cat_ids = p.catIds if p.useCats else [-1]
        self.ious = self._compute_all_oks([(i, c) for i in p.imgIds for c in cat_ids])
        max_det = p.maxDets[-1]
        self.evalImgs = [self.evaluateImg(i, c, rng, max_det) for c in cat_ids for rng in p.areaRng for i in p.imgIds]
        self._paramsEval = copy.deepcopy(p)
        print("DONE.")

    def evaluateImg(self, imgId, catId, aRng, maxDet):
        p = self.params
        gt, dt = self._gts[imgId, catId], self._dts[imgId, catId]
        if not gt and not dt:
            return None
        for g in gt:
            g["_ignore"] = 1 if (g["ignore"] or g["area"] < aRng[0] or g["area"] > aRng[1]) else 0
        gtind = np.argsort([g["_ignore"] for g in gt], kind="mergesort")
        gt = [gt[i] for i in gtind]
        dtind = np.argsort([-d["score"] for d in dt], kind="mergesort")
        dt = [dt[i] for i in dtind[:maxDet]]
        iscrowd = [int(g.get("iscrowd", 0)) for g in gt]
        ious = self.ious[imgId, catId]
        ious = ious[:, gtind] if len(ious) > 0 else ious
        T, G, D = len(p.iouThrs), len(gt), len(dt)
        gtm, dtm = np.zeros((T, G)), np.zeros((T, D))
        gt_ig = np.array([g["_ignore"] for g in gt])
        dt_ig = np.zeros((T, D))
        if len(ious) != 0:
            for ti, t in enumerate(p.iouThrs):
                for di, d in enumerate(dt):
                    best, m = min([t, 1 - 1e-10]), -1
                    for gi in range(G):
                        if gtm[ti, gi] > 0 and not iscrowd[gi]:
                            continue                      # already claimed
                        if m > -1 and gt_ig[m] == 0 and gt_ig[gi] == 1:
                            break                         # a regular match exists; only ignored ones follow
                        if ious[di, gi] < best:
                            continue
                        best, m = ious[di, gi], gi
                    if m == -1:
                        continue
                    dt_ig[ti, di] = gt_ig[m]
                    dtm[ti, di] = gt[m]["id"]
                    gtm[ti, m] = d["id"]
        out_of_range = np.array([d["area"] < aRng[0] or d["area"] > aRng[1] for d in dt]).reshape((1, len(dt)))
        dt_ig = np.logical_or(dt_ig, np.logical_and(dtm == 0, np.repeat(out_of_range, T, 0)))
        return {"image_id": imgId, "category_id": catId, "aRng": aRng, "maxDet": maxDet,
                "dtIds": [d["id"] for d in dt], "gtIds": [g["id"] for g in gt], "dtMatches": dtm, "gtMatches": gtm,
                "dtScores": [d["score"] for d in dt], "gtIgnore": gt_ig, "dtIgnore": dt_ig}

    # ------------------------------------------------------------------ accumulate
    def accumulate(self, p=None):
        print("Accumulating evaluation results...")
        if not self.evalImgs:
            print("Please run evaluate() first")
        if p is None:
            p = self.params
        p.catIds = p.catIds if p.useCats == 1 else [-1]
        T, R, K, A, M = len(p.iouThrs), len(p.recThrs), len(p.catIds) if p.useCats else 1, len(p.areaRng), len(p.maxDets)
        precision = -np.ones((T, R, K, A, M))
        recall = -np.ones((T, K, A, M))
        scores = -np.ones((T, R, K, A, M))
        pe = self._paramsEval
        cat_ids = pe.catIds if pe.useCats else [-1]
        set_k, set_m, set_i = set(cat_ids), set(pe.maxDets), set(pe.imgIds)
        set_a = set(map(tuple, pe.areaRng))
        k_list = [n for n, k in enumerate(p.catIds) if k in set_k]
        m_list = [m for m in p.maxDets if m in set_m]
        a_list = [n for n, a in enumerate(map(tuple, p.areaRng)) if a in set_a]
        i_list = [n for n, i in enumerate(p.imgIds) if i in set_i]
        I0, A0 = len(pe.imgIds), len(pe.areaRng)
        for k, k0 in enumerate(k_list):
            for a, a0 in enumerate(a_list):
                for m, max_det in enumerate(m_list):
                    E = [self.evalImgs[k0 * A0 * I0 + a0 * I0 + i] for i in i_list]
                    E = [e for e in E if e is not None]
                    if not E:
                        continue
                    dt_scores = np.concatenate([e["dtScores"][0:max_det] for e in E])
                    order = np.argsort(-dt_scores, kind="mergesort")
                    sorted_scores = dt_scores[order]
                    dtm = np.concatenate([e["dtMatches"][:, 0:max_det] for e in E], axis=1)[:, order]
                    dt_ig = np.concatenate([e["dtIgnore"][:, 0:max_det] for e in E], axis=1)[:, order]
                    gt_ig = np.concatenate([e["gtIgnore"] for e in E])
                    npig = np.count_nonzero(gt_ig == 0)
                    if npig == 0:
                        continue
                    tps = np.logical_and(dtm, np.logical_not(dt_ig))
                    fps = np.logical_and(np.logical_not(dtm), np.logical_not(dt_ig))
                    tp_sum = np.cumsum(tps, axis=1).astype(dtype=float)
                    fp_sum = np.cumsum(fps, axis=1).astype(dtype=float)
                    for t, (tp, fp) in enumerate(zip(tp_sum, fp_sum)):
                        nd = len(tp)
                        rc = tp / npig
                        pr = (tp / (fp + tp + np.spacing(1))).tolist()
                        recall[t, k, a, m] = rc[-1] if nd else 0
                        for i in range(nd - 1, 0, -1):     # monotone (non-increasing in recall) envelope
                            if pr[i] > pr[i - 1]:
                                pr[i - 1] = pr[i]
                        q, ss = np.zeros(R), np.zeros(R)
                        for ri, pi in enumerate(np.searchsorted(rc, p.recThrs, side="left")):
                            if pi >= nd:
                                break                      # recall level never reached: precision stays 0
                            q[ri], ss[ri] = pr[pi], sorted_scores[pi]
                        precision[t, :, k, a, m] = q
                        scores[t, :, k, a, m] = ss
        self.eval = {"params": p, "counts": [T, R, K, A, M], "precision": precision, "recall": recall, "scores": scores}
        print("DONE.")

    # ------------------------------------------------------------------ summarize
    def _summarize(self, ap=1, iouThr=None, areaRng="all", maxDets=20):
        p = self.params
        line = " {:<18} {} @[ IoU={:<9} | area={:>6s} | maxDets={:>3d} ] = {:0.3f}"
        title = "Average Precision" if ap == 1 else "Average Recall"
        kind = "(AP)" if ap == 1 else "(AR)"
        iou = "{:0.2f}:{:0.2f}".format(p.iouThrs[0], p.iouThrs[-1]) if iouThr is None else "{:0.2f}".format(iouThr)
        aind = [i for i, lbl in enumerate(p.areaRngLbl) if lbl == areaRng]
        mind = [i for i, md in enumerate(p.maxDets) if md == maxDets]
        s = self.eval["precision"] if ap == 1 else self.eval["recall"]
        if iouThr is not None:
            s = s[np.where(iouThr == p.iouThrs)[0]]
        s = s[:, :, :, aind, mind] if ap == 1 else s[:, :, aind, mind]
        mean_s = -1 if len(s[s > -1]) == 0 else np.mean(s[s > -1])
        print(line.format(title, kind, iou, areaRng, maxDets, mean_s))
        return mean_s

    def summarize(self):
        if not self.eval:
            raise Exception("Please run accumulate() first")
        stats = np.zeros((10,))
        stats[0] = self._summarize(1, maxDets=20)
        stats[1] = self._summarize(1, maxDets=20, iouThr=.5)
        stats[2] = self._summarize(1, maxDets=20, iouThr=.75)
        stats[3] = self._summarize(1, maxDets=20, areaRng="medium")
        stats[4] = self._summarize(1, maxDets=20, areaRng="large")
        stats[5] = self._summarize(0, maxDets=20)
        stats[6] = self._summarize(0, maxDets=20, iouThr=.5)
        stats[7] = self._summarize(0, maxDets=20, iouThr=.75)
        stats[8] = self._summarize(0, maxDets=20, areaRng="medium")
        stats[9] = self._summarize(0, maxDets=20, areaRng="large")
        self.stats = stats

    def __str__(self):
        self.summarize()
        return ""
