"""Hot-path part of the reference's dataset_builder.py: keypoint scaling (:107-111) and target-heatmap
rendering (:220-238), on the GPU.  TFRecord parsing / JPEG decode / imgaug augmentation are out of
scope; `SyntheticDatasetBuilder` provides the (images, heatmaps) contract that
Trainer consumes, with targets rendered on device from synthetic keypoints.
"""
from __future__ import annotations

import numpy as np

from . import _lib, ops


def scale_keypoints(kps, extent, label_extent):
    """dataset_builder.py:107-111: `kps /= float32(extent); kps *= label_extent` (two float32 ops)."""
    kps = np.asarray(kps, dtype=np.float32)
    return ((kps / np.float32(extent)).astype(np.float32) * np.float32(label_extent)).astype(np.float32)


def np_gen_heatmaps(kps_x, kps_y, kps_v, label_shape=(64, 64, 17)):
    """One sample, host arrays in / host array out -- same contract as DatasetBuilder.np_gen_heatmaps."""
    h, w, k = label_shape
    if not (len(kps_x) == len(kps_y) == k):
        raise AssertionError("expected one (x, y) per keypoint")
    out = ops.render_targets(np.asarray(kps_x, np.float32)[None], np.asarray(kps_y, np.float32)[None],
                             np.asarray(kps_v)[None], h, w)
    return out[0].cpu().numpy()


class SyntheticDatasetBuilder:
    """Stands in for DatasetBuilder(config, ratio): `.build_datasets()` -> (ds_train, ds_valid), infinite
    iterables of (images (B,256,256,3) f32 CUDA, heatmaps (B,64,64,17) f32 CUDA); `.num_train_examples`,
    `.num_valid_examples`.  Targets come from hgb_render_targets inside the iterator (config 2 of BASELINE.json)."""

    def __init__(self, config, num_train_examples=64, num_valid_examples=32, seed=0):
        self.image_shape = tuple(config.IMAGE_SHAPE)
        self.label_shape = tuple(config.LABEL_SHAPE)
        self.num_keypoints = int(config.NUM_KEYPOINTS)
        self.batch_size = int(config.BATCH_SIZE)
        self.num_train_examples, self.num_valid_examples = int(num_train_examples), int(num_valid_examples)
        self.seed = seed

    def _stream(self, seed):
        torch = _lib.require_cuda()
        gen = torch.Generator(device="cuda").manual_seed(seed)
        h, w, k = self.label_shape
        while True:
            img = torch.rand((self.batch_size,) + self.image_shape, device="cuda", generator=gen)
            kx = torch.rand((self.batch_size, k), device="cuda", generator=gen) * (w + 8) - 4
            ky = torch.rand((self.batch_size, k), device="cuda", generator=gen) * (h + 8) - 4
            kv = torch.randint(0, 3, (self.batch_size, k), device="cuda", generator=gen, dtype=torch.int32)
            yield img, ops.render_targets(kx, ky, kv, h, w)

    def build_datasets(self):
        return self._stream(self.seed), self._stream(self.seed + 1)

    def np_gen_heatmaps(self, kps_x, kps_y, kps_v):
        return np_gen_heatmaps(kps_x, kps_y, kps_v, self.label_shape)
