"""Drop-in for the reference's gen_tfrecords.py (the writer of the dataset DatasetBuilder reads): one tf.train.Example per
annotated person -- square box, crop + pad, keypoints relative to the box with out-of-box / unlabelled joints zeroed
(gen_tfrecords.py:12-86) -- written as TFRecord shards named `file_{train|valid}_NN-COUNT.tfrec` (:88-115) without
TensorFlow.  The crop runs on the GPU (`hgb_crop_resize` at the crop's own size = an exact copy); JPEG *encoding* is host
library work (cv2 / libjpeg, as tf.io.encode_jpeg is) and can be replaced through `encode_fn`.
"""
from __future__ import annotations

import os

import numpy as np

from . import tfrecord
from .utilities import data_utils


def filter_keypoints(keypoints, bbox):
    """gen_tfrecords.py:31-59: COCO (x, y, v) triplets -> xs, ys, vs relative to the square box; a joint survives iff it lies
    strictly inside the box and v > 0, everything else becomes (0, 0, 0)."""
    xs, ys, vs = [], [], []
    for i in range(0, len(keypoints), 3):
        x, y, v = keypoints[i] - bbox[0], keypoints[i + 1] - bbox[1], int(keypoints[i + 2])
        if 0 < x < bbox[2] and 0 < y < bbox[3] and v > 0:
            xs.append(x)
            ys.append(y)
            vs.append(v)
        else:
            xs.append(0)
            ys.append(0)
            vs.append(0)
    return xs, ys, vs


def _device_crop_u8(image, bbox):
    """crop_and_pad on the device, back as uint8: for uint8 input the kernel computes float32(u) * float32(1/255), and
    rint(that * 255) == u for every u in 0..255 (tests/test_cpu_tfrecord.py), so the copy is exact."""
    import torch
    crop = data_utils.crop_and_pad(image, bbox)
    if np.asarray(image).dtype == np.uint8:
        return torch.round(crop * 255.0).to(torch.uint8).cpu().numpy()
    return crop.cpu().numpy()


def _encode_jpeg(rgb_u8):
    import cv2
    ok, enc = cv2.imencode(".jpg", cv2.cvtColor(rgb_u8, cv2.COLOR_RGB2BGR), [cv2.IMWRITE_JPEG_QUALITY, 95])   # tf.io.encode_jpeg default quality
    if not ok:
        raise RuntimeError("JPEG encoding failed")
    return enc.tobytes()


def example_features(image, image_path, example, index, bbox_scale, crop_fn=None, encode_fn=None):
    """The feature dictionary of create_example (gen_tfrecords.py:70-85) as {name: value} ready for tfrecord.build_example."""
    bbox = data_utils.transform_bbox_square(example["bbox"], scale=bbox_scale)
    crop = (crop_fn or _device_crop_u8)(image, bbox)
    xs, ys, vs = filter_keypoints(example["keypoints"], bbox)
    return {
        "ann_id": np.array([int(example["ann_id"])], np.int64), "image_id": np.array([int(index)], np.int64),
        "image": (encode_fn or _encode_jpeg)(crop), "image_path": image_path.encode(), "coco_url": example["coco_url"].encode(),
        "width": np.array([int(crop.shape[1])], np.int64), "height": np.array([int(crop.shape[0])], np.int64),
        "keypoints/x": np.asarray(xs, np.float32), "keypoints/y": np.asarray(ys, np.float32), "keypoints/vis": np.asarray(vs, np.int64),
        "keypoints/num": np.array([sum(v > 0 for v in vs)], np.int64),
        "bbox_x": np.array([bbox[0]], np.float32), "bbox_y": np.array([bbox[1]], np.float32),
        "original_bbox": np.asarray(example["bbox"], np.float32),
    }


def create_example(image, image_path, example, index, bbox_scale, crop_fn=None, encode_fn=None):
    """gen_tfrecords.py:12-86 -> the serialized tf.train.Example."""
    return tfrecord.build_example(example_features(image, image_path, example, index, bbox_scale, crop_fn, encode_fn))


def gen_TFRecords(df, config, is_train, read_image=None):
    """gen_tfrecords.py:88-115: `df` is the reference's merged COCO dataframe (index = image id, columns image_path, bbox,
    keypoints, ann_id, coco_url); NUM_EXAMPLER_PER_TFRECORD examples per shard."""
    per = config.NUM_EXAMPLER_PER_TFRECORD
    shards = len(df) // per + (1 if len(df) % per else 0)
    folder = config.TRAIN_TFRECORDS_DIR if is_train else config.VALID_TFRECORDS_DIR
    images_dir = config.TRAIN_IMAGES_DIR if is_train else config.VALID_IMAGES_DIR
    os.makedirs(folder, exist_ok=True)
    if read_image is None:
        import cv2

        def read_image(path):
            return cv2.cvtColor(cv2.imread(path, cv2.IMREAD_COLOR), cv2.COLOR_BGR2RGB)
    for shard in range(shards):
        examples = df[shard * per:(shard + 1) * per]
        name = folder + "/file_" + folder.split("/")[-1] + "_%.2i-%i.tfrec" % (shard, len(examples))
        payloads = []
        for index, row in examples.iterrows():
            path = os.path.join(images_dir, row["image_path"])
            payloads.append(create_example(read_image(path), path, row, index, config.BBOX_SCALE))
        tfrecord.write_records(name, payloads)
    print("TFRecords generated at", folder)
