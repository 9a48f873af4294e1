"""The reference's dataset_builder.py on the GPU: keypoint scaling (:107-111), target-heatmap rendering (:220-238),
and the per-example work of the input pipeline -- resize (:99), flip + affine augmentation of image and keypoints
(:143-185, :270-300) and the colour augmentation (:190-204) -- as batched kernels (`make_train_label_batch`).
The random draws stay on the host (a handful of scalars per example, numpy Generator instead of imgaug / tf.random);
everything that touches pixels runs in libhgb200.  `SyntheticDatasetBuilder` provides the (images, heatmaps)
contract Trainer consumes without TFRecords.
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


# ------------------------------------------------------------------ augmentation (dataset_builder.py:143-204)
def flip_partner(num_keypoints, index_flip_pairs):
    """Permutation applied by flip_labels (dataset_builder.py:270-300): slot k receives joint partner[k]."""
    partner = list(range(num_keypoints))
    for a, b in index_flip_pairs:
        partner[a], partner[b] = partner[b], partner[a]
    return np.asarray(partner, np.int32)


def affine_matrix(height, width, scale, rotate_deg, shift_add):
    """Forward 3x3 float64 matrix of imgaug's Affine(scale, rotate): about (size/2 - shift_add); shift_add is 0.5 for
    pixel arrays and 0 for keypoint coordinates (imgaug `to_matrix` / `to_matrix_cba`)."""
    sy, sx = height / 2.0 - shift_add, width / 2.0 - shift_add
    rot = np.deg2rad(rotate_deg)
    lin = np.array([[scale * np.cos(rot), -scale * np.sin(rot), 0.0], [scale * np.sin(rot), scale * np.cos(rot), 0.0], [0.0, 0.0, 1.0]])
    to_topleft = np.array([[1.0, 0.0, -sx], [0.0, 1.0, -sy], [0.0, 0.0, 1.0]])
    to_center = np.array([[1.0, 0.0, sx], [0.0, 1.0, sy], [0.0, 0.0, 1.0]])
    return to_center @ (lin @ to_topleft)


def _opencv_inverse(forward):
    """The inverse map cv2.warpAffine derives from a forward 2x3 matrix (its order of double operations)."""
    m = np.array(forward[:2], dtype=np.float64).reshape(-1).copy()
    det = m[0] * m[4] - m[1] * m[3]
    det = 1.0 / det if det != 0 else 0.0
    a11, a22 = m[4] * det, m[0] * det
    m[0], m[4] = a11, a22
    m[1] *= -det
    m[3] *= -det
    b1 = -m[0] * m[2] - m[1] * m[5]
    b2 = -m[3] * m[2] - m[4] * m[5]
    m[2], m[5] = b1, b2
    return m.reshape(2, 3)


def augment_1_batch(images, kps_x, kps_y, kps_v, flip, scale, rotate_deg, label_shape=(64, 64, 17),
                    index_flip_pairs=((1, 2), (3, 4), (5, 6), (7, 8), (9, 10), (11, 12), (13, 14), (15, 16))):
    """DatasetBuilder.np_augment_1 for a batch with the random draws given: flip (N,) bool, scale (N,), rotate_deg (N,).
    images (N,H,W,3) float32 -> warped images; keypoints in label-map coordinates -> augmented (N,K) float32 x, y."""
    images_shape = tuple(images.shape)
    n, h, w = images_shape[0], images_shape[1], images_shape[2]
    lh, lw, k = label_shape
    inv = np.stack([_opencv_inverse(affine_matrix(h, w, float(s), float(r), 0.5)) for s, r in zip(scale, rotate_deg)]) \
        if n else np.zeros((0, 2, 3))
    fwd = np.stack([affine_matrix(lh, lw, float(s), float(r), 0.0)[:2] for s, r in zip(scale, rotate_deg)]) if n else np.zeros((0, 2, 3))
    aug_images = ops.augment_affine(images, inv, flip)
    ax, ay = ops.augment_keypoints(kps_x, kps_y, kps_v, flip, fwd, flip_partner(k, index_flip_pairs), lw)
    return aug_images, ax, ay


def augment_2_batch(images, brightness_delta, contrast_factor, saturation_factor, hue_delta):
    """DatasetBuilder.augment_2 (:190-204) in place on a CUDA batch, the four tf.image.random_* draws given per example."""
    params = np.stack([np.asarray(brightness_delta, np.float32), np.asarray(contrast_factor, np.float32),
                       np.asarray(saturation_factor, np.float32), np.asarray(hue_delta, np.float32)], axis=1)
    return ops.color_augment(images, params)


def draw_augmentation(rng, n):
    """One set of random draws per example with the reference's ranges: flip p=.5 (:160), scale U(.75,1.25) and rotation
    U(-30,30) degrees (:167), brightness +-0.2, contrast U(.5,2), saturation U(.75,1.25), hue +-0.1 (:194-197)."""
    return {"flip": rng.integers(0, 2, n).astype(bool), "scale": rng.uniform(0.75, 1.25, n), "rotate_deg": rng.uniform(-30.0, 30.0, n),
            "brightness_delta": rng.uniform(-0.2, 0.2, n), "contrast_factor": rng.uniform(0.5, 2.0, n),
            "saturation_factor": rng.uniform(0.75, 1.25, n), "hue_delta": rng.uniform(-0.1, 0.1, n)}


def make_train_label_batch(images, kps_x, kps_y, kps_v, draws, label_shape=(64, 64, 17)):
    """DatasetBuilder.make_train_label (:68-78) for a batch: augmentation 1, augmentation 2, target rendering.  As in the
    reference the renderer receives the ORIGINAL visibility flags (the swapped copy made by flip_labels is dropped at :185)."""
    aug, ax, ay = augment_1_batch(images, kps_x, kps_y, kps_v, draws["flip"], draws["scale"], draws["rotate_deg"], label_shape)
    augment_2_batch(aug, draws["brightness_delta"], draws["contrast_factor"], draws["saturation_factor"], draws["hue_delta"])
    return aug, ops.render_targets(ax, ay, kps_v, label_shape[0], label_shape[1])


def make_valid_label_batch(images, kps_x, kps_y, kps_v, label_shape=(64, 64, 17)):
    """DatasetBuilder.make_valid_label (:81-85)."""
    return images, ops.render_targets(kps_x, kps_y, kps_v, label_shape[0], label_shape[1])
