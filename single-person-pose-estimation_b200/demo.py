"""Drop-in for the compute path of the reference's demo.py (Demo.detect, demo.py:25-71): detector boxes -> square boxes ->
crop / pad / resize -> hourglass -> heat-map decode, for every person of a frame in one batch on the GPU:

    frame (uint8 or float)  --one launch-->  (N,256,256,3) crops     data_utils.crop_and_resize  (demo.py:44-52)
    crops                   --network--->    last stack's heat maps   HourglassModel.forward_device (demo.py:57)
    heat maps               --one launch-->  (N,17,3) keypoints       hgb_decode v2               (demo.py:58-63)

The frame is uploaded once; only N*17*3 floats come back.  Plotting (Demo.show / show_bboxes / show_separate /
create_overlay) is matplotlib UI code and out of scope; the attributes those methods read are populated identically.
"""
from __future__ import annotations

import numpy as np

from . import _lib
from .utilities import data_utils


def _boxes_from_detector_result(result, person_conf_thres):
    """demo.py:28-38: YOLOv5 result -> [(xmin, ymin, w, h)] of the 'person' rows.  Also accepts an (n,4) xyxy array."""
    if hasattr(result, "pandas"):
        df = result.pandas().xyxy[0]
        df = df[(df["name"] == "person") & (df["confidence"] > person_conf_thres)]
        cols = [df[c].values for c in ("xmin", "ymin", "xmax", "ymax")]
    else:
        arr = np.asarray(result, dtype=np.float64).reshape(-1, 4)
        cols = [arr[:, i] for i in range(4)]
    return [(xmin, ymin, xmax - xmin, ymax - ymin) for xmin, ymin, xmax, ymax in zip(*cols)]


class Demo:
    def __init__(self, person_detector, keypoints_detetor, cfg, max_num_ppl=6, person_conf_thres=1e-6, keypoints_conf_thres=1e-6):
        self.person_detector = person_detector
        self.keypoints_detetor = keypoints_detetor
        self.person_conf_thres = person_conf_thres
        self.keypoints_conf_thres = keypoints_conf_thres
        self.COCO_SKELETON = getattr(cfg, "COCO_SKELETON", None)
        self.max_num_ppl = max_num_ppl
        self.cfg = cfg

    def detect(self, image):
        cfg = self.cfg
        bboxes = _boxes_from_detector_result(self.person_detector(image), self.person_conf_thres)[: self.max_num_ppl]
        transformed_bboxes = [data_utils.transform_bbox_square(b, cfg.BBOX_SCALE) for b in bboxes]
        keypoints_list, cropped = [], []
        if bboxes:
            torch = _lib.require_cuda()
            crops = data_utils.crop_and_resize(image, transformed_bboxes, cfg.IMAGE_HEIGHT, cfg.IMAGE_WIDTH)
            model = self.keypoints_detetor
            if hasattr(model, "forward_device"):
                last = model.forward_device(crops, training=False)[-1]
                _idx, kps = data_utils.decode_batch(last, self.keypoints_conf_thres, 2)
                kps = kps.cpu().numpy()
            else:                                             # any object with the Keras predict contract
                last = model.predict(crops.cpu().numpy())[-1]
                kps = np.stack([data_utils.heatmaps_to_keypoints_v2(h, self.keypoints_conf_thres) for h in last])
            kps[:, :, 0] /= cfg.LABEL_WIDTH
            kps[:, :, 1] /= cfg.LABEL_HEIGHT
            keypoints_list = [k for k in kps]
            cropped = [c for c in crops]
            del torch
        self.image = image
        self.cropped_images = cropped
        self.original_bboxes = bboxes
        self.square_bboxes = transformed_bboxes
        self.keypoints_list = keypoints_list
        return keypoints_list

    def keypoints_in_image(self):
        """Keypoints in frame coordinates: x * box_w + box_x, y * box_h + box_y (the mapping Demo.show draws, demo.py:86-90);
        undetected joints (0, 0, 0) stay zero."""
        out = []
        for kps, box in zip(self.keypoints_list, self.square_bboxes):
            k = np.array(kps, dtype=np.float64)
            found = (k[:, 0] != 0) & (k[:, 1] != 0)
            k[:, 0] = np.where(found, k[:, 0] * box[2] + box[0], 0.0)
            k[:, 1] = np.where(found, k[:, 1] * box[3] + box[1], 0.0)
            out.append(k)
        return out

    def _no_ui(self, *_a, **_k):
        raise NotImplementedError("plotting is out of scope of the B200 path (SURVEY.md section 8f); use keypoints_in_image()")

    show = show_bboxes = show_separate = create_overlay = _no_ui
