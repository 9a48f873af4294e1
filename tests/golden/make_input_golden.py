"""Generate tests/golden/input_golden.npz for the input path.

Run in the build container only (needs /root/reference and cv2):
    python tests/golden/make_input_golden.py

What is recorded, and from what:
  * crop_and_pad / transform_bbox_square: the REFERENCE's own functions (utilities/data_utils.py:16-98), executed with a
    numpy stand-in for the three TensorFlow calls they make (tf.shape, tf.image.pad_to_bounding_box,
    tf.image.crop_to_bounding_box -- zero padding / slicing with TF's argument checks);
  * DatasetBuilder.flip_labels: the REFERENCE's own static method (dataset_builder.py:270-300) with a minimal stand-in
    for imgaug's Keypoint / KeypointsOnImage containers (no arithmetic of its own);
  * warp: the real cv2.warpAffine(INTER_LINEAR, BORDER_CONSTANT, 0) on float32 images -- the call imgaug's Affine makes.
TensorFlow's resize / colour kernels and imgaug's matrix construction cannot be executed here (packages absent).
"""
import os
import sys
import types

import cv2
import numpy as np

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(OUT)))


class _Keypoint:
    def __init__(self, x, y):
        self.x, self.y = x, y


class _KeypointsOnImage:
    def __init__(self, keypoints, shape):
        self.keypoints, self.shape = keypoints, shape

    def to_xy_array(self):
        return np.array([[k.x, k.y] for k in self.keypoints], dtype=np.float32).reshape(-1, 2)


def _pad_to_bounding_box(image, offset_height, offset_width, target_height, target_width):
    h, w, c = image.shape
    after_h, after_w = target_height - offset_height - h, target_width - offset_width - w
    if offset_height < 0 or offset_width < 0 or after_h < 0 or after_w < 0:
        raise ValueError("pad_to_bounding_box: bad arguments")
    out = np.zeros((target_height, target_width, c), image.dtype)
    out[offset_height:offset_height + h, offset_width:offset_width + w] = image
    return out


def _crop_to_bounding_box(image, offset_height, offset_width, target_height, target_width):
    h, w, _ = image.shape
    if offset_height < 0 or offset_width < 0 or target_height <= 0 or target_width <= 0:
        raise ValueError("crop_to_bounding_box: bad arguments")
    if w < target_width + offset_width:
        raise ValueError("width must be >= target + offset.")
    if h < target_height + offset_height:
        raise ValueError("height must be >= target + offset.")
    return image[offset_height:offset_height + target_height, offset_width:offset_width + target_width]


def import_reference():
    def stub(name, **attrs):
        m = types.ModuleType(name)
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[name] = m
        return m

    tf = stub("tensorflow")
    tf.data = types.SimpleNamespace(Dataset=object)
    tf.shape = lambda t: np.array(t.shape)
    tf.image = types.SimpleNamespace(pad_to_bounding_box=_pad_to_bounding_box, crop_to_bounding_box=_crop_to_bounding_box)
    stub("matplotlib")
    stub("matplotlib.pyplot")
    stub("matplotlib.patches")
    stub("pycocotools")
    stub("pycocotools.coco", COCO=object)
    stub("pycocotools.cocoeval", COCOeval=object)
    stub("imgaug")
    stub("imgaug.augmenters")
    stub("imgaug.augmentables", Keypoint=_Keypoint, KeypointsOnImage=_KeypointsOnImage)
    sys.path.insert(0, REF)
    import dataset_builder  # noqa
    from utilities import data_utils  # noqa
    return data_utils, dataset_builder


def main():
    data_utils, dataset_builder = import_reference()
    rng = np.random.default_rng(20)
    out = {}

    # ---- crop_and_pad on a 90x120 uint8 frame: interior, each border, corners, scaled detector boxes
    frame = rng.integers(0, 256, (90, 120, 3), dtype=np.uint8)
    det_boxes = [(30.2, 20.7, 40.5, 50.1), (-5.5, 10.0, 50.0, 40.0), (80.3, 40.9, 60.0, 70.0), (2.0, -12.7, 30.3, 30.3),
                 (60.0, 5.0, 59.9, 20.0), (-20.4, -30.6, 160.2, 150.8), (0.0, 0.0, 120.0, 90.0), (100.5, 70.5, 19.5, 19.5),
                 (10.0, 10.0, 5.9, 80.0)]
    squares, params, crops = [], [], []
    for b in det_boxes:
        for scale in (1, 1.25):
            sq = data_utils.transform_bbox_square(b, scale)
            crop = data_utils.crop_and_pad(frame, sq)
            squares.append(sq)
            crops.append(crop)
    out["frame"] = frame
    out["det_boxes"] = np.array(det_boxes, np.float64)
    out["square_boxes"] = np.array(squares, np.float64)
    out["crop_shapes"] = np.array([c.shape[:2] for c in crops], np.int64)
    out["crops_flat"] = np.concatenate([c.reshape(-1) for c in crops])

    # ---- flip_labels (reference static method)
    pairs = [[1, 2], [3, 4], [5, 6], [7, 8], [9, 10], [11, 12], [13, 14], [15, 16]]
    xs = rng.random(17).astype(np.float32) * 64
    ys = rng.random(17).astype(np.float32) * 64
    vs = rng.integers(0, 3, 17).astype(np.int64)
    kpsoi = _KeypointsOnImage([_Keypoint(x, y) for x, y in zip(xs, ys)], shape=(64, 64, 3))
    flipped, fvs = dataset_builder.DatasetBuilder.flip_labels(kpsoi, pairs, vs.copy())
    out["flip_in_xyv"] = np.stack([xs, ys, vs.astype(np.float32)], 1)
    out["flip_out_xy"] = flipped.to_xy_array()
    out["flip_out_v"] = np.asarray(fvs)

    # ---- cv2.warpAffine on float32 images, matrices as imgaug builds them (restated; see oracle/input_oracle.py)
    from oracle import input_oracle as iorc
    img = rng.random((64, 64, 3), dtype=np.float32)
    cases = [(1.0, 0.0), (0.75, -30.0), (1.25, 30.0), (0.9, 17.3), (1.1, -5.5), (1.0, 90.0)]
    warped, mats = [], []
    for sc, rot in cases:
        m = iorc.affine_matrix(64, 64, sc, rot, 0.5)[:2]
        mats.append(m)
        warped.append(cv2.warpAffine(img, m, dsize=(64, 64), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_CONSTANT, borderValue=0))
    out["warp_image"] = img
    out["warp_cases"] = np.array(cases, np.float64)
    out["warp_mats"] = np.array(mats, np.float64)
    out["warp_out"] = np.stack(warped)
    # cv2.resize uses the sampling grid tf.image.resize(bilinear) uses (half-pixel centres, no antialias); rounding order differs
    big = rng.random((75, 50, 3), dtype=np.float32)
    out["resize_in"] = big
    out["resize_cv2_64"] = cv2.resize(big, (64, 64), interpolation=cv2.INTER_LINEAR)

    np.savez_compressed(os.path.join(OUT, "input_golden.npz"), **out)
    print("wrote input_golden.npz:", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
