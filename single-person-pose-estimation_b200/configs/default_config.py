"""Constants the hot path reads from the reference's `cfg` object (configs/default_config.py).
Any object exposing these names (the reference module included) can be passed wherever `config` is taken."""
import os

import numpy as np

NUM_KEYPOINTS = 17
MIN_NUM_KEYPOINTS = 5
NUM_EXAMPLER_PER_TFRECORD = 2048
IMAGE_WIDTH, IMAGE_HEIGHT = 256, 256
IMAGE_SHAPE = (IMAGE_HEIGHT, IMAGE_WIDTH, 3)
LABEL_WIDTH, LABEL_HEIGHT = 64, 64
LABEL_SHAPE = (LABEL_HEIGHT, LABEL_WIDTH, NUM_KEYPOINTS)
GAUSSIAN_KERNEL = 7
HM_SIGMA = 1
HM_ACTIVATION = "sigmoid"

HG_NUM_CHANNELS = 256
HG_NUM_STACKS = 2

BATCH_SIZE = 16
SHUFFLE_BUFFER = 1000
LEARNING_RATE = 0.01
BBOX_SCALE = 1.25

DATASET_DIR = "dataset"
IMAGES_DIR = os.path.join(DATASET_DIR, "images")
TRAIN_IMAGES_DIR, VALID_IMAGES_DIR = os.path.join(IMAGES_DIR, "train2017"), os.path.join(IMAGES_DIR, "val2017")
ANNOT_DIR = os.path.join(DATASET_DIR, "annotations")
TRAIN_ANNOT_FILE = os.path.join(ANNOT_DIR, "person_keypoints_train2017.json")
VALID_ANNOT_FILE = os.path.join(ANNOT_DIR, "person_keypoints_val2017.json")
TFRECORDS_DIR = os.path.join(DATASET_DIR, "tfrecords")
TRAIN_TFRECORDS_DIR, VALID_TFRECORDS_DIR = os.path.join(TFRECORDS_DIR, "train"), os.path.join(TFRECORDS_DIR, "valid")

TEMPORARY_DIR = "temp"
CHECKPOINTS_PATH = os.path.join(TEMPORARY_DIR, "checkpoints")
LOGS_PATH = os.path.join(TEMPORARY_DIR, "logs")

COCO_KEYPOINT_LABELS = ["nose", "left_eye", "right_eye", "left_ear", "right_ear", "left_shoulder", "right_shoulder",
                        "left_elbow", "right_elbow", "left_wrist", "right_wrist", "left_hip", "right_hip",
                        "left_knee", "right_knee", "left_ankle", "right_ankle"]
COCO_INDEX_FLIP_PAIRS = [[i, i + 1] for i in range(1, 17, 2)]
# limb list, 0-based joint indices (COCO person skeleton)
COCO_SKELETON = np.array([[16, 14], [14, 12], [17, 15], [15, 13], [12, 13], [6, 12], [7, 13], [6, 7], [6, 8], [7, 9], [8, 10], [9, 11],
                          [2, 3], [1, 2], [1, 3], [2, 4], [3, 5], [4, 6], [5, 7]]) - 1
