"""Input-path oracle (oracle/input_oracle.py) against the committed fixtures produced by the reference's own functions and
by cv2 (tests/golden/make_input_golden.py), and the product's host bookkeeping against the oracle's.  CPU only."""
import os

import numpy as np
import pytest

from oracle import input_oracle as iorc

PAIRS = [[1, 2], [3, 4], [5, 6], [7, 8], [9, 10], [11, 12], [13, 14], [15, 16]]


@pytest.fixture(scope="module")
def g(golden_dir):
    return np.load(os.path.join(golden_dir, "input_golden.npz"))


def _crops(g):
    at = 0
    for shape in g["crop_shapes"]:
        n = int(shape[0]) * int(shape[1]) * 3
        yield g["crops_flat"][at:at + n].reshape(int(shape[0]), int(shape[1]), 3)
        at += n


def test_crop_and_pad_matches_reference(g):
    import hgb200
    frame = g["frame"]
    for sq, want in zip(g["square_boxes"], _crops(g)):
        got = iorc.crop_and_pad(frame, tuple(sq))
        assert got.shape == want.shape
        np.testing.assert_array_equal(got, want)
        # the product's host bookkeeping is the same integer arithmetic
        assert hgb200.data_utils.crop_and_pad_params(frame.shape[0], frame.shape[1], tuple(sq)) == \
            iorc.crop_and_pad_params(frame.shape[0], frame.shape[1], tuple(sq))


def test_square_bbox_matches_reference(g):
    import hgb200
    k = 0
    for b in g["det_boxes"]:
        for scale in (1, 1.25):
            assert tuple(g["square_boxes"][k]) == hgb200.data_utils.transform_bbox_square(tuple(b), scale)
            k += 1


def test_square_bbox_known_answers_from_the_reference_notebook():
    """dev/test_tfrecords.ipynb cells 15 and 39: the printed results of the reference's transform_bbox_square."""
    import hgb200
    assert hgb200.data_utils.transform_bbox_square([603.15, 125.6, 36.85, 66.16]) == (588.4949999999999, 125.60000000000001, 66.16, 66.16)
    assert hgb200.data_utils.transform_bbox_square([163.73, 126.42, 265.69, 480.4], 1.25) == (-3.6750000000000114, 66.37, 600.5, 600.5)


def test_crop_window_rejected_like_tensorflow():
    with pytest.raises(ValueError):
        iorc.crop_and_pad_params(90, 120, (10.0, 10.0, 0.5, 0.5))


def test_flip_labels_matches_reference(g):
    xyv = g["flip_in_xyv"]
    v = xyv[:, 2].astype(np.int64)
    # identity affine: only the flip's label swap remains (x mirrored separately by Fliplr in the reference)
    x, y = iorc.augment_keypoints(64 - xyv[:, 0], xyv[:, 1], np.ones(17, np.int64), True, 1.0, 0.0, 64, 64, PAIRS)
    np.testing.assert_allclose(x, g["flip_out_xy"][:, 0], rtol=0, atol=1e-5)
    np.testing.assert_array_equal(y, g["flip_out_xy"][:, 1])
    swapped = v.copy()
    for a, b in PAIRS:
        swapped[a], swapped[b] = swapped[b], swapped[a]
    np.testing.assert_array_equal(swapped, g["flip_out_v"])


def test_warp_affine_is_bit_exact_against_cv2(g):
    import hgb200
    for (scale, rot), mat, want in zip(g["warp_cases"], g["warp_mats"], g["warp_out"]):
        m = iorc.affine_matrix(64, 64, scale, rot, 0.5)[:2]
        np.testing.assert_array_equal(m, mat)
        np.testing.assert_array_equal(iorc.warp_affine(g["warp_image"], m), want)
        np.testing.assert_array_equal(hgb200.dataset_builder.affine_matrix(64, 64, scale, rot, 0.5)[:2], mat)
        np.testing.assert_array_equal(hgb200.dataset_builder._opencv_inverse(mat), iorc.invert_affine_cv(mat))


def test_resize_uses_the_half_pixel_grid(g):
    got = iorc.resize_bilinear(g["resize_in"], 64, 64)
    np.testing.assert_allclose(got, g["resize_cv2_64"], rtol=0, atol=3e-7)      # same grid as cv2; rounding order differs
    same = iorc.resize_bilinear(g["resize_in"], 75, 50)
    np.testing.assert_array_equal(same, g["resize_in"])                         # identity at the native size


def test_color_augment_properties():
    rng = np.random.default_rng(0)
    img = rng.random((32, 32, 3), dtype=np.float32)
    out = iorc.color_augment(img, 0.1, 1.3, 1.1, 0.05)
    assert out.min() == 0.0 and out.max() == 1.0
    # neutral draws leave hue/saturation untouched up to rounding: result is the min-max normalised input
    neutral = iorc.color_augment(img, 0.0, 1.0, 1.0, 0.0)
    want = (img - img.min()) / (img.max() - img.min())
    np.testing.assert_allclose(neutral, want, atol=2e-6)
    # a hue rotation by a full turn is the identity
    np.testing.assert_allclose(iorc.color_augment(img, 0.0, 1.0, 1.0, 1.0), neutral, atol=2e-6)
    # grey stays grey under saturation / hue changes
    grey = np.repeat(rng.random((8, 8, 1), dtype=np.float32), 3, axis=2)
    og = iorc.color_augment(grey, 0.0, 1.0, 1.2, 0.07)
    assert np.abs(og[..., 0] - og[..., 1]).max() < 1e-6 and np.abs(og[..., 1] - og[..., 2]).max() < 1e-6


def test_hue_and_saturation_agree_with_colorsys():
    """Independent check of the restated TF kernels against the standard library's HSV round trip."""
    import colorsys
    rng = np.random.default_rng(5)
    px = rng.random((200, 3)).astype(np.float32)
    r, g, b = iorc._adjust_hue(px[:, 0], px[:, 1], px[:, 2], 0.083)
    h, s, v = iorc._rgb_to_hsv(px[:, 0], px[:, 1], px[:, 2])
    r2, g2, b2 = iorc._hsv_to_rgb(h, np.minimum(np.float32(1), s * np.float32(1.2)), v)
    for i in range(len(px)):
        hh, ss, vv = colorsys.rgb_to_hsv(*[float(c) for c in px[i]])
        want = colorsys.hsv_to_rgb((hh + 0.083) % 1.0, ss, vv)
        np.testing.assert_allclose([r[i], g[i], b[i]], want, atol=2e-6)
        want2 = colorsys.hsv_to_rgb(hh, min(1.0, ss * 1.2), vv)
        np.testing.assert_allclose([r2[i], g2[i], b2[i]], want2, atol=2e-6)


def test_product_matrices_equal_the_oracle_on_random_draws():
    """The host bookkeeping of augment_1_batch (forward matrices for image / keypoints, OpenCV's inverse) is bit-identical
    to the oracle's for the whole range of the reference's draws."""
    import hgb200
    rng = np.random.default_rng(7)
    for _ in range(500):
        scale, rot = rng.uniform(0.75, 1.25), rng.uniform(-30, 30)
        for h, w, shift in ((256, 256, 0.5), (64, 64, 0.0), (75, 50, 0.5)):
            want = iorc.affine_matrix(h, w, scale, rot, shift)
            got = hgb200.dataset_builder.affine_matrix(h, w, scale, rot, shift)
            np.testing.assert_array_equal(got, want)
            np.testing.assert_array_equal(hgb200.dataset_builder._opencv_inverse(got[:2]), iorc.invert_affine_cv(want[:2]))
