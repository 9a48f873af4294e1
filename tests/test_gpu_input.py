"""Input-path kernels (csrc/input_kernels.cu) through the C ABI against the oracle and the committed fixtures:
crop_and_pad / resize (demo.py:44-50, dataset_builder.py:99), flip + affine warp of image and keypoints
(dataset_builder.py:143-185), colour augmentation (:190-204).  Bit-exact wherever the reference arithmetic is
deterministic; the colour path within 2e-6 (the per-channel mean is a reduction whose order differs)."""
import os

import numpy as np
import pytest

from oracle import heatmap_oracle as horc
from oracle import input_oracle as iorc

pytestmark = pytest.mark.gpu
PAIRS = [[1, 2], [3, 4], [5, 6], [7, 8], [9, 10], [11, 12], [13, 14], [15, 16]]


@pytest.fixture(scope="module")
def hgb():
    import hgb200
    return hgb200


@pytest.fixture(scope="module")
def g(golden_dir):
    return np.load(os.path.join(golden_dir, "input_golden.npz"))


def _crops(g):
    at = 0
    for shape in g["crop_shapes"]:
        n = int(shape[0]) * int(shape[1]) * 3
        yield g["crops_flat"][at:at + n].reshape(int(shape[0]), int(shape[1]), 3)
        at += n


def test_crop_and_pad_matches_reference_fixture(hgb, g):
    frame = g["frame"]
    for sq, want in zip(g["square_boxes"], _crops(g)):
        got = hgb.data_utils.crop_and_pad(frame, tuple(sq)).cpu().numpy()
        np.testing.assert_array_equal(got, iorc.convert_u8(want))            # the reference crops the uint8 frame's float copy


def test_crop_and_resize_batch_bit_exact(hgb, g):
    frame = g["frame"]
    boxes = [tuple(b) for b in g["square_boxes"]]
    for src in (frame, iorc.convert_u8(frame)):
        got = hgb.data_utils.crop_and_resize(src, boxes, 256, 256).cpu().numpy()
        assert got.shape == (len(boxes), 256, 256, 3)
        for n, b in enumerate(boxes):
            np.testing.assert_array_equal(got[n], iorc.crop_resize(src, b, 256, 256))
    assert hgb.data_utils.crop_and_resize(frame, [], 256, 256).shape == (0, 256, 256, 3)
    with pytest.raises(ValueError):
        hgb.data_utils.crop_and_resize(frame, [(10.0, 10.0, 0.5, 0.5)])


def test_resize_ragged_batch_down_and_up(hgb, g):
    rng = np.random.default_rng(1)
    imgs = [rng.integers(0, 256, (h, w, 3), dtype=np.uint8) for h, w in ((480, 640), (333, 97), (64, 64), (1, 1), (1024, 700))]
    got = hgb.data_utils.resize_images(imgs, 256, 256).cpu().numpy()
    for n, im in enumerate(imgs):
        np.testing.assert_array_equal(got[n], iorc.crop_resize(im, None, 256, 256))
    f = g["resize_in"]
    np.testing.assert_allclose(hgb.data_utils.resize_images([f], 64, 64).cpu().numpy()[0], g["resize_cv2_64"], rtol=0, atol=3e-7)


def test_affine_warp_bit_exact_against_cv2_fixture(hgb, g):
    img = g["warp_image"]
    n = len(g["warp_cases"])
    inv = np.stack([hgb.dataset_builder._opencv_inverse(m) for m in g["warp_mats"]])
    got = hgb.ops.augment_affine(np.repeat(img[None], n, 0), inv, np.zeros(n, bool)).cpu().numpy()
    np.testing.assert_array_equal(got, g["warp_out"])
    flipped = hgb.ops.augment_affine(np.repeat(img[None], n, 0), inv, np.ones(n, bool)).cpu().numpy()
    for k, (scale, rot) in enumerate(g["warp_cases"]):
        np.testing.assert_array_equal(flipped[k], iorc.augment_image(img, True, scale, rot))


def test_augment_1_batch_matches_oracle_at_full_size(hgb):
    rng = np.random.default_rng(2)
    n = 12
    images = rng.random((n, 256, 256, 3), dtype=np.float32)
    kx = (rng.random((n, 17)) * 72 - 4).astype(np.float32)
    ky = (rng.random((n, 17)) * 72 - 4).astype(np.float32)
    kv = rng.integers(0, 3, (n, 17)).astype(np.int64)
    d = hgb.dataset_builder.draw_augmentation(np.random.default_rng(3), n)
    d["flip"][:2] = [True, False]
    aug, ax, ay = hgb.dataset_builder.augment_1_batch(images, kx, ky, kv, d["flip"], d["scale"], d["rotate_deg"])
    aug, ax, ay = aug.cpu().numpy(), ax.cpu().numpy(), ay.cpu().numpy()
    for i in range(n):
        np.testing.assert_array_equal(aug[i], iorc.augment_image(images[i], bool(d["flip"][i]), d["scale"][i], d["rotate_deg"][i]))
        wx, wy = iorc.augment_keypoints(kx[i], ky[i], kv[i], bool(d["flip"][i]), d["scale"][i], d["rotate_deg"][i], 64, 64, PAIRS)
        np.testing.assert_array_equal(ax[i], wx)
        np.testing.assert_array_equal(ay[i], wy)


def test_flip_labels_matches_reference_fixture(hgb, g):
    xyv = g["flip_in_xyv"]
    ident = np.array([[[1.0, 0.0, 0.0], [0.0, 1.0, 0.0]]])
    partner = hgb.dataset_builder.flip_partner(17, PAIRS)
    ox, oy = hgb.ops.augment_keypoints((64 - xyv[:, 0])[None], xyv[None, :, 1], np.ones((1, 17), np.int32), [True], ident, partner, 64)
    np.testing.assert_allclose(ox.cpu().numpy()[0], g["flip_out_xy"][:, 0], rtol=0, atol=1e-5)
    np.testing.assert_array_equal(oy.cpu().numpy()[0], g["flip_out_xy"][:, 1])


def test_color_augment_matches_oracle(hgb):
    import torch
    rng = np.random.default_rng(4)
    n = 6
    images = rng.random((n, 256, 256, 3), dtype=np.float32)
    images[1] *= 0.2                                                     # dark image: negative values after contrast
    images[2, :, :, :] = images[2, :, :, :1]                             # grey image: zero chroma everywhere
    d = hgb.dataset_builder.draw_augmentation(np.random.default_rng(5), n)
    x = torch.as_tensor(images).cuda()
    out = hgb.dataset_builder.augment_2_batch(x, d["brightness_delta"], d["contrast_factor"], d["saturation_factor"], d["hue_delta"])
    assert out.data_ptr() == x.data_ptr()
    got = out.cpu().numpy()
    for i in range(n):
        want = iorc.color_augment(images[i], d["brightness_delta"][i], d["contrast_factor"][i], d["saturation_factor"][i], d["hue_delta"][i])
        assert got[i].min() == 0.0 and got[i].max() == 1.0
        np.testing.assert_allclose(got[i], want, rtol=0, atol=2e-6)
        assert (got[i] != want).mean() < 0.02                            # all but a few pixels are bit-identical


def test_train_label_batch_properties_at_baseline_size(hgb):
    """Batch 256 at 256x256 (BASELINE config sizes): size-independent properties."""
    import torch
    n = 256
    gen = torch.Generator(device="cuda").manual_seed(0)
    images = torch.rand((n, 256, 256, 3), device="cuda", generator=gen)
    kx = torch.rand((n, 17), device="cuda", generator=gen) * 72 - 4
    ky = torch.rand((n, 17), device="cuda", generator=gen) * 72 - 4
    kv = torch.randint(0, 3, (n, 17), device="cuda", generator=gen, dtype=torch.int32)
    ones, zeros = np.ones(n), np.zeros(n)
    # identity draw: image unchanged, keypoints unchanged where visible
    same, ax, ay = hgb.dataset_builder.augment_1_batch(images, kx, ky, kv, zeros.astype(bool), ones, zeros)
    assert torch.equal(same, images)
    assert torch.equal(ax, torch.where(kv > 0, kx, torch.zeros_like(kx)))
    # two flips restore the image exactly and swap every label pair back
    once, fx, fy = hgb.dataset_builder.augment_1_batch(images, kx, ky, kv, ones.astype(bool), ones, zeros)
    twice, _, _ = hgb.dataset_builder.augment_1_batch(once, kx, ky, kv, ones.astype(bool), ones, zeros)
    assert torch.equal(twice, images) and torch.equal(once, images.flip(2))
    assert torch.equal(fy[:, 1], torch.where(kv[:, 2] > 0, ky[:, 2], torch.zeros_like(ky[:, 2])))
    # the full training label: normalised images, targets identical to rendering the augmented keypoints
    d = hgb.dataset_builder.draw_augmentation(np.random.default_rng(6), n)
    aug, heat = hgb.dataset_builder.make_train_label_batch(images, kx, ky, kv, d)
    assert aug.shape == (n, 256, 256, 3) and heat.shape == (n, 64, 64, 17)
    assert float(aug.amin()) == 0.0 and float(aug.amax()) == 1.0
    assert torch.equal(aug.amin(dim=(1, 2, 3)), torch.zeros(n, device="cuda")) and torch.equal(aug.amax(dim=(1, 2, 3)), torch.ones(n, device="cuda"))
    _, ax, ay = hgb.dataset_builder.augment_1_batch(images, kx, ky, kv, d["flip"], d["scale"], d["rotate_deg"])
    want = horc.render_targets(ax[:4].cpu().numpy(), ay[:4].cpu().numpy(), kv[:4].cpu().numpy(), 64, 64)
    np.testing.assert_array_equal(heat[:4].cpu().numpy(), want)
