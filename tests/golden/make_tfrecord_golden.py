"""Generate tests/golden/tfrecord_golden.npz: the feature dictionaries the REFERENCE's own `create_example`
(gen_tfrecords.py:12-86) produces, executed here with recording stand-ins for the TensorFlow calls it makes
(tf.shape, tf.image.pad/crop_to_bounding_box through crop_and_pad, tf.io.encode_jpeg -> the raw crop bytes,
tf.train.Example / Features / Feature / *List -> plain containers).  Pins the dataset writer's host logic: square box,
crop, keypoint filtering, feature names and value types.  Run in the build container only:
    python tests/golden/make_tfrecord_golden.py"""
import json
import os
import sys
import types

import numpy as np

OUT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, OUT)
import make_input_golden as mig  # noqa: E402  (reuses its numpy pad/crop stand-ins and module stubs)


class _List:
    def __init__(self, value):
        self.value = list(value)


class _Feature:
    def __init__(self, bytes_list=None, float_list=None, int64_list=None):
        self.kind = "bytes" if bytes_list is not None else "float" if float_list is not None else "int64"
        self.value = (bytes_list or float_list or int64_list).value


class _Features:
    def __init__(self, feature):
        self.feature = feature


class _Example:
    def __init__(self, features):
        self.features = features


class _Encoded:
    def __init__(self, arr):
        self.arr = np.asarray(arr)

    def numpy(self):
        return b"RAW" + self.arr.astype(np.uint8).tobytes()


def main():
    mig.import_reference()
    tf = sys.modules["tensorflow"]
    tf.train = types.SimpleNamespace(Example=_Example, Features=_Features, Feature=_Feature, BytesList=_List, FloatList=_List, Int64List=_List)
    tf.io = types.SimpleNamespace(encode_jpeg=_Encoded)
    stub = types.ModuleType("coco_df")
    stub.gen_trainval_df = None
    sys.modules["coco_df"] = stub
    cfg = types.ModuleType("configs.default_config")
    pkg = types.ModuleType("configs")
    pkg.default_config = cfg
    sys.modules["configs"], sys.modules["configs.default_config"] = pkg, cfg
    import gen_tfrecords  # noqa: E402

    rng = np.random.default_rng(30)
    image = rng.integers(0, 256, (150, 200, 3), dtype=np.uint8)
    rows = []
    boxes = [[40.0, 30.0, 60.0, 90.0], [-5.0, 20.0, 80.0, 40.0], [120.5, 60.25, 90.0, 100.0], [10.0, -8.0, 30.0, 30.0]]
    for i, bbox in enumerate(boxes):
        kps = []
        for k in range(17):
            inside = rng.random() < 0.7
            x = rng.uniform(bbox[0] - 20, bbox[0] + bbox[2] + 20) if not inside else rng.uniform(bbox[0] + 1, bbox[0] + bbox[2] - 1)
            y = rng.uniform(bbox[1] - 20, bbox[1] + bbox[3] + 20) if not inside else rng.uniform(bbox[1] + 1, bbox[1] + bbox[3] - 1)
            kps += [float(np.floor(x)), float(np.floor(y)), int(rng.integers(0, 3))]
        rows.append({"bbox": bbox, "keypoints": kps, "ann_id": 900000 + i, "coco_url": f"http://images.cocodataset.org/val2017/{i:012d}.jpg"})
    out = {"image": image, "rows": np.array(json.dumps(rows))}
    for i, row in enumerate(rows):
        for scale in (1.25, 1):
            ex = gen_tfrecords.create_example(image, f"dataset/images/val2017/{i}.jpg", row, 5000 + i, scale)
            for name, feat in ex.features.feature.items():
                key = f"ex{i}_{scale}_{name}"
                if feat.kind == "bytes":
                    out[key] = np.frombuffer(feat.value[0], dtype=np.uint8)
                elif feat.kind == "float":
                    out[key] = np.array(feat.value, dtype=np.float64)
                else:
                    out[key] = np.array(feat.value, dtype=np.int64)
    np.savez_compressed(os.path.join(OUT, "tfrecord_golden.npz"), **out)
    print("wrote tfrecord_golden.npz with", len(out), "arrays")


if __name__ == "__main__":
    main()
