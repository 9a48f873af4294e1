"""B200-native stacked-hourglass heatmap path (drop-in for the reference's Python API).

The package mirrors the reference layout for the hot path only:
    model/hourglass.py   create_hourglass_model(...)
    loss.py              weighted_mse / IOU / weighed_keypoint_mse / mean_squared_error
    trainer.py           Trainer.train() / resume_training()
    utilities/data_utils.py  heatmaps_to_keypoints_v1/_v2
    eval.py              predict_ds / eval_PCK / eval_OKS   (cocoeval.py: the COCO AP protocol behind eval_OKS)
    dataset_builder.py   DatasetBuilder (TFRecords -> device batches), augmentation, target rendering   (tfrecord.py: formats)
    demo.py              Demo.detect (boxes -> crops -> keypoints)
Everything numerical runs in libhgb200.so (hand-written sm_100a CUDA) through ctypes;
torch tensors only hold device memory.  There is no CPU fallback: importing works
without a GPU (so the ABI can be inspected), calling an op without one raises.
"""
from . import _lib  # noqa: F401  (loads libhgb200.so; raises if it is missing)

__all__ = ["_lib"]
from . import ops  # noqa: E402,F401

__all__ += ["ops"]
from . import loss, parallel  # noqa: E402,F401
from .model import hourglass  # noqa: E402,F401
from .model.hourglass import Adam, HourglassModel, create_hourglass_model  # noqa: E402,F401

__all__ += ["loss", "parallel", "hourglass", "Adam", "HourglassModel", "create_hourglass_model"]
from . import callbacks, cocoeval, dataset_builder, demo, eval, gen_tfrecords, tfrecord, trainer  # noqa: E402,F401
from .configs import default_config  # noqa: E402,F401
from .trainer import Trainer  # noqa: E402,F401
from .utilities import data_utils  # noqa: E402,F401
from .utilities.data_utils import heatmaps_to_keypoints_v1, heatmaps_to_keypoints_v2  # noqa: E402,F401

__all__ += ["callbacks", "cocoeval", "demo", "gen_tfrecords", "tfrecord", "dataset_builder", "eval", "trainer", "default_config", "Trainer", "data_utils",
            "heatmaps_to_keypoints_v1", "heatmaps_to_keypoints_v2"]
