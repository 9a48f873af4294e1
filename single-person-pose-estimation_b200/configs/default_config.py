"""Constants the hot path reads from the reference's `cfg` object (configs/default_config.py).
Any object exposing these names (the reference module included) can be passed wherever `config` is taken."""
import os

NUM_KEYPOINTS = 17
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
LEARNING_RATE = 0.01
BBOX_SCALE = 1.25

TEMPORARY_DIR = "temp"
CHECKPOINTS_PATH = os.path.join(TEMPORARY_DIR, "checkpoints")
LOGS_PATH = os.path.join(TEMPORARY_DIR, "logs")

COCO_KEYPOINT_LABELS = ["nose", "left_eye", "right_eye", "left_ear", "right_ear", "left_shoulder", "right_shoulder",
                        "left_elbow", "right_elbow", "left_wrist", "right_wrist", "left_hip", "right_hip",
                        "left_knee", "right_knee", "left_ankle", "right_ankle"]
COCO_INDEX_FLIP_PAIRS = [[i, i + 1] for i in range(1, 17, 2)]
