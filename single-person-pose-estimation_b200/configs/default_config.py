"""The constant names the hot path reads from the reference's `cfg` object (configs/default_config.py), with the
reference's values.  Any object exposing these names -- the reference module itself included -- can be passed wherever a
`config` is taken; nothing in the package imports this module implicitly except as a default."""
from os.path import join as _join

import numpy as np

# ---- geometry: 256x256 RGB crops in, 64x64 heat maps for the 17 COCO person joints out
NUM_KEYPOINTS = 17
IMAGE_HEIGHT = IMAGE_WIDTH = 256
LABEL_HEIGHT = LABEL_WIDTH = 64
IMAGE_SHAPE = (IMAGE_HEIGHT, IMAGE_WIDTH, 3)
LABEL_SHAPE = (LABEL_HEIGHT, LABEL_WIDTH, NUM_KEYPOINTS)
GAUSSIAN_KERNEL, HM_SIGMA = 7, 1                 # stored by DatasetBuilder, the renderer's 7x7 sigma-1 stamp is fixed
HM_ACTIVATION = "sigmoid"
BBOX_SCALE = 1.25                                # person box -> square crop, enlarged

# ---- network / optimisation defaults
HG_NUM_STACKS, HG_NUM_CHANNELS = 2, 256
BATCH_SIZE, LEARNING_RATE = 16, 0.01
SHUFFLE_BUFFER = 1000

# ---- dataset generation
MIN_NUM_KEYPOINTS = 5                            # people with fewer labelled joints are dropped (coco_df.gen_trainval_df)
NUM_EXAMPLER_PER_TFRECORD = 2048                 # [sic] the reference's spelling

# ---- where things live
DATASET_DIR = "dataset"
IMAGES_DIR, ANNOT_DIR, TFRECORDS_DIR = (_join(DATASET_DIR, d) for d in ("images", "annotations", "tfrecords"))
TRAIN_IMAGES_DIR, VALID_IMAGES_DIR = _join(IMAGES_DIR, "train2017"), _join(IMAGES_DIR, "val2017")
TRAIN_ANNOT_FILE, VALID_ANNOT_FILE = (_join(ANNOT_DIR, f"person_keypoints_{s}2017.json") for s in ("train", "val"))
TRAIN_TFRECORDS_DIR, VALID_TFRECORDS_DIR = _join(TFRECORDS_DIR, "train"), _join(TFRECORDS_DIR, "valid")
TEMPORARY_DIR = "temp"
CHECKPOINTS_PATH, LOGS_PATH = _join(TEMPORARY_DIR, "checkpoints"), _join(TEMPORARY_DIR, "logs")

# ---- COCO person skeleton: nose, then left/right pairs from the head down
_PAIRED = ("eye", "ear", "shoulder", "elbow", "wrist", "hip", "knee", "ankle")
COCO_KEYPOINT_LABELS = ["nose"] + [f"{side}_{part}" for part in _PAIRED for side in ("left", "right")]
COCO_INDEX_FLIP_PAIRS = [[2 * k + 1, 2 * k + 2] for k in range(len(_PAIRED))]           # (left, right) joint indices
_LIMBS_1_BASED = ((16, 14), (14, 12), (17, 15), (15, 13), (12, 13), (6, 12), (7, 13), (6, 7), (6, 8), (7, 9), (8, 10), (9, 11),
                  (2, 3), (1, 2), (1, 3), (2, 4), (3, 5), (4, 6), (5, 7))
COCO_SKELETON = np.array(_LIMBS_1_BASED) - 1
# matplotlib colour per joint (only the reference's plotting helpers read it)
COCO_KEYPOINT_COLORS = "red brown chocolate orange tan lime teal navy violet black coral yellow gold cyan green orchid indigo".split()
