"""CPU restatement of the reference's INPUT path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module; the product path never does.

What it restates, and how each piece is pinned:

  convert_u8             tf.image.convert_image_dtype(uint8 -> float32)            demo.py:44, dataset_builder.py:264
  crop_and_pad_params    utilities/data_utils.py:48-98 (integer pad/crop bookkeeping)  -- pure Python in the reference; the
                         restatement below is checked against the reference function itself run with a numpy stand-in for
                         tf.image.pad_to_bounding_box / crop_to_bounding_box (tests/golden/make_input_golden.py)
  crop_and_pad           the same, producing the cropped image
  resize_bilinear        tf.image.resize(..., bilinear) = ResizeBilinear with half-pixel centres, no antialias
                         (dataset_builder.py:99,133, demo.py:50).  TensorFlow is absent: **parity unpinned against live TF**;
                         the arithmetic is the published kernel (lower/upper/lerp, top/bottom lerp in float32) and is
                         cross-checked against cv2.resize(INTER_LINEAR), which uses the same sampling grid.
  affine_matrix          imgaug 0.4 `Affine(scale, rotate)` matrix about the array centre (images: -0.5 shift) and
                         `Fliplr` (dataset_builder.py:163-172).  imgaug is absent: **parity unpinned**, restated from its
                         published source.
  warp_affine            cv2.warpAffine(INTER_LINEAR, BORDER_CONSTANT 0) on float32 images -- what imgaug calls.  cv2 IS
                         present in the build image: pinned bit-for-bit by tests/golden/input_golden.npz.
  augment_keypoints      Fliplr + flip_labels + Affine on keypoints, and the final visibility filter
                         (dataset_builder.py:143-185, 270-300)
  color_augment          tf.image.adjust_brightness / adjust_contrast / adjust_saturation / adjust_hue + min-max
                         normalisation (dataset_builder.py:190-204) with the random draws passed in.  **Unpinned** (TF absent);
                         restated from the published CPU kernels.
"""
from __future__ import annotations

import numpy as np

F = np.float32


# ------------------------------------------------------------------ dtype conversion / crop / resize
def convert_u8(image_u8: np.ndarray) -> np.ndarray:
    return (image_u8.astype(F) * F(1.0 / 255)).astype(F)


def crop_and_pad_params(image_height: int, image_width: int, square_bbox):
    """(x0, y0, crop_w, crop_h): crop pixel (cy, cx) reads source pixel (cy + y0, cx + x0), zero outside the source.
    Raises ValueError where tf.image.crop_to_bounding_box would (crop window larger than the padded image)."""
    x, y, w, h = square_bbox
    xmin, ymin, xmax, ymax = x, y, x + w, y + h
    offset_width = offset_height = 0
    target_width, target_height = image_width, image_height
    if xmin < 0:
        offset_width = int(abs(x))
        target_width += offset_width
    if ymin < 0:
        offset_height = int(abs(y))
        target_height += offset_height
    if xmax > image_width:
        target_width += int(xmax - image_width) + 1
    if ymax > image_height:
        target_height += int(ymax - image_height) + 1
    cy, cx, ch, cw = int(max(ymin, 0)), int(max(xmin, 0)), int(h), int(w)
    if cw <= 0 or ch <= 0:
        raise ValueError("target_width and target_height must be > 0")
    if target_width < cw + cx:
        raise ValueError("width must be >= target + offset.")
    if target_height < ch + cy:
        raise ValueError("height must be >= target + offset.")
    return cx - offset_width, cy - offset_height, cw, ch


def crop_and_pad(image: np.ndarray, square_bbox) -> np.ndarray:
    x0, y0, cw, ch = crop_and_pad_params(image.shape[0], image.shape[1], square_bbox)
    out = np.zeros((ch, cw, image.shape[2]), image.dtype)
    ys, xs = np.arange(ch) + y0, np.arange(cw) + x0
    vy, vx = (ys >= 0) & (ys < image.shape[0]), (xs >= 0) & (xs < image.shape[1])
    out[np.ix_(vy, vx)] = image[np.ix_(ys[vy], xs[vx])]
    return out


def _interp_weights(out_size: int, in_size: int):
    scale = F(in_size) / F(out_size)
    pos = ((np.arange(out_size).astype(F) + F(0.5)) * scale).astype(F) - F(0.5)
    fl = np.floor(pos)
    lower = np.maximum(fl.astype(np.int64), 0)
    upper = np.minimum(np.ceil(pos).astype(np.int64), in_size - 1)
    return lower, upper, (pos - fl).astype(F)


def resize_bilinear(image: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """float32 (H,W,C) -> (out_h,out_w,C), every operation rounded to float32 in the kernel's order."""
    image = image.astype(F)
    ylo, yhi, yl = _interp_weights(out_h, image.shape[0])
    xlo, xhi, xl = _interp_weights(out_w, image.shape[1])
    xl = xl[None, :, None]
    yl = yl[:, None, None]
    tl, tr = image[ylo][:, xlo], image[ylo][:, xhi]
    bl, br = image[yhi][:, xlo], image[yhi][:, xhi]
    top = (tl + ((tr - tl).astype(F) * xl).astype(F)).astype(F)
    bot = (bl + ((br - bl).astype(F) * xl).astype(F)).astype(F)
    return (top + ((bot - top).astype(F) * yl).astype(F)).astype(F)


def crop_resize(image: np.ndarray, square_bbox, out_h: int, out_w: int) -> np.ndarray:
    """demo.py:44-50: convert (if uint8), crop_and_pad, resize."""
    if image.dtype == np.uint8:
        image = convert_u8(image)
    return resize_bilinear(crop_and_pad(image, square_bbox) if square_bbox is not None else image, out_h, out_w)


# ------------------------------------------------------------------ augmentation 1: flip + affine
def affine_matrix(height: int, width: int, scale: float, rotate_deg: float, shift_add: float) -> np.ndarray:
    """3x3 float64 forward matrix: translate(-c) -> scale+rotate -> translate(+c), c = size/2 - shift_add
    (shift_add 0.5 for images, 0 for keypoints)."""
    sy, sx = height / 2.0 - shift_add, width / 2.0 - shift_add
    rot = np.deg2rad(rotate_deg)
    a = np.array([[scale * np.cos(rot), -scale * np.sin(rot), 0.0],
                  [scale * np.sin(rot), scale * np.cos(rot), 0.0],
                  [0.0, 0.0, 1.0]])
    to_topleft = np.array([[1.0, 0.0, -sx], [0.0, 1.0, -sy], [0.0, 0.0, 1.0]])
    to_center = np.array([[1.0, 0.0, sx], [0.0, 1.0, sy], [0.0, 0.0, 1.0]])
    return to_center @ (a @ to_topleft)


def invert_affine_cv(m: np.ndarray) -> np.ndarray:
    """cv2.warpAffine's own inversion of the forward 2x3 matrix (double precision, its operation order)."""
    m = np.array(m[:2], dtype=np.float64).reshape(-1).copy()
    d = m[0] * m[4] - m[1] * m[3]
    d = 1.0 / d if d != 0 else 0.0
    a11, a22 = m[4] * d, m[0] * d
    m[0] = a11
    m[1] *= -d
    m[3] *= -d
    m[4] = a22
    b1 = -m[0] * m[2] - m[1] * m[5]
    b2 = -m[3] * m[2] - m[4] * m[5]
    m[2], m[5] = b1, b2
    return m.reshape(2, 3)


def warp_coords(inv: np.ndarray, height: int, width: int):
    """Fixed-point source coordinates of cv2.warpAffine: integer pixel (sx, sy) and 5-bit fractions (fx, fy)."""
    AB, ROUND = 1024, 16
    xs = np.arange(width, dtype=np.float64)
    ys = np.arange(height, dtype=np.float64)
    adelta = np.rint(inv[0, 0] * xs * AB).astype(np.int64)
    bdelta = np.rint(inv[1, 0] * xs * AB).astype(np.int64)
    x0 = np.rint((inv[0, 1] * ys + inv[0, 2]) * AB).astype(np.int64) + ROUND
    y0 = np.rint((inv[1, 1] * ys + inv[1, 2]) * AB).astype(np.int64) + ROUND
    X = (x0[:, None] + adelta[None, :]) >> 5
    Y = (y0[:, None] + bdelta[None, :]) >> 5
    sx = np.clip(X >> 5, -32768, 32767)
    sy = np.clip(Y >> 5, -32768, 32767)
    return sx, sy, X & 31, Y & 31


def warp_affine(image: np.ndarray, forward_2x3: np.ndarray) -> np.ndarray:
    """cv2.warpAffine(image f32, M, (W,H), flags=INTER_LINEAR, borderMode=BORDER_CONSTANT, borderValue=0)."""
    image = image.astype(F)
    H, W = image.shape[:2]
    sx, sy, fx, fy = warp_coords(invert_affine_cv(forward_2x3), H, W)
    t = (np.arange(32, dtype=F) * F(1.0 / 32)).astype(F)
    w0, w1 = (F(1.0) - t).astype(F), t
    wx0, wx1, wy0, wy1 = w0[fx], w1[fx], w0[fy], w1[fy]

    def tap(yy, xx):
        ok = (yy >= 0) & (yy < H) & (xx >= 0) & (xx < W)
        v = image[np.clip(yy, 0, H - 1), np.clip(xx, 0, W - 1)]
        return np.where(ok[..., None], v, F(0))

    c = lambda a: a[..., None]  # noqa: E731
    acc = (tap(sy, sx) * c((wy0 * wx0).astype(F))).astype(F)
    acc = (acc + (tap(sy, sx + 1) * c((wy0 * wx1).astype(F))).astype(F)).astype(F)
    acc = (acc + (tap(sy + 1, sx) * c((wy1 * wx0).astype(F))).astype(F)).astype(F)
    acc = (acc + (tap(sy + 1, sx + 1) * c((wy1 * wx1).astype(F))).astype(F)).astype(F)
    return acc


def augment_keypoints(kps_x, kps_y, kps_v, flip: bool, scale: float, rotate_deg: float, label_h: int, label_w: int, flip_pairs):
    """dataset_builder.py:143-185: keypoints with v <= 0 become (0,0) first; optional Fliplr (x -> W - x) with left/right
    label swap of x, y AND v; Affine about (W/2, H/2); coordinates of joints whose (swapped) v <= 0 are zeroed.
    Returns float32 (K,), (K,) -- the visibility array handed to the heat-map renderer afterwards is the ORIGINAL one
    (dataset_builder.py:78-82 passes kps_v, not the swapped copy)."""
    v = np.array(kps_v).copy()
    x = np.where(v > 0, np.asarray(kps_x, F), F(0)).astype(F)
    y = np.where(v > 0, np.asarray(kps_y, F), F(0)).astype(F)
    if flip:
        x = (F(label_w) - x).astype(F)
        for a, b in flip_pairs:
            x[a], x[b] = x[b], x[a]
            y[a], y[b] = y[b], y[a]
            v[a], v[b] = v[b], v[a]
    m = affine_matrix(label_h, label_w, scale, rotate_deg, 0.0)
    xd, yd = x.astype(np.float64), y.astype(np.float64)
    xa = m[0, 0] * xd + m[0, 1] * yd + m[0, 2]
    ya = m[1, 0] * xd + m[1, 1] * yd + m[1, 2]
    return np.where(v > 0, xa.astype(F), F(0)).astype(F), np.where(v > 0, ya.astype(F), F(0)).astype(F)


def augment_image(image: np.ndarray, flip: bool, scale: float, rotate_deg: float) -> np.ndarray:
    if flip:
        image = image[:, ::-1]
    m = affine_matrix(image.shape[0], image.shape[1], scale, rotate_deg, 0.5)
    return warp_affine(np.ascontiguousarray(image), m[:2])


# ------------------------------------------------------------------ augmentation 2: colour
def _rgb_to_hsv(r, g, b):
    vv = np.maximum(r, np.maximum(g, b))
    rng = (vv - np.minimum(r, np.minimum(g, b))).astype(F)
    with np.errstate(divide="ignore", invalid="ignore"):
        s = np.where(vv > 0, (rng / vv).astype(F), F(0)).astype(F)
        norm = (F(1.0) / (F(6.0) * rng).astype(F)).astype(F)
        hr = (norm * (g - b).astype(F)).astype(F)
        hg = ((norm * (b - r).astype(F)).astype(F) + F(2.0 / 6.0)).astype(F)
        hb = ((norm * (r - g).astype(F)).astype(F) + F(4.0 / 6.0)).astype(F)
    hh = np.where(r == vv, hr, np.where(g == vv, hg, hb))
    hh = np.where(rng <= 0, F(0), hh)
    hh = np.where(hh < 0, (hh + F(1)).astype(F), hh).astype(F)
    return hh, s, vv


def _hsv_to_rgb(h, s, v):
    c = (s * v).astype(F)
    m = (v - c).astype(F)
    dh = (h * F(6)).astype(F)
    cat = dh.astype(np.int32)
    fm = dh.copy()
    for _ in range(4):
        fm = np.where(fm <= 0, (fm + F(2)).astype(F), fm)
        fm = np.where(fm >= 2, (fm - F(2)).astype(F), fm)
    x = (c * (F(1) - np.abs((fm - F(1)).astype(F))).astype(F)).astype(F)
    z = np.zeros_like(c)
    rr = np.select([cat == 0, cat == 1, cat == 2, cat == 3, cat == 4, cat == 5], [c, x, z, z, x, c], z)
    gg = np.select([cat == 0, cat == 1, cat == 2, cat == 3, cat == 4, cat == 5], [x, c, c, x, z, z], z)
    bb = np.select([cat == 0, cat == 1, cat == 2, cat == 3, cat == 4, cat == 5], [z, z, x, c, c, x], z)
    return (rr + m).astype(F), (gg + m).astype(F), (bb + m).astype(F)


def _adjust_hue(r, g, b, delta):
    """Published CPU kernel of tf.image.adjust_hue: hue in [0,6) sextants with the (v_min, v_max) range kept."""
    r, g, b = r.astype(F), g.astype(F), b.astype(F)
    c1 = (r < g) & (b < r)
    c3 = (r < g) & ~(b < r) & (b > g)
    c2 = (r < g) & ~(b < r) & ~(b > g)
    c0 = ~(r < g) & (b < g)
    c4 = ~(r < g) & ~(b < g) & (b > r)
    vmax = np.select([c1, c3, c2, c0, c4], [g, b, g, r, b], r)
    vmid = np.select([c1, c3, c2, c0, c4], [r, g, b, g, r], b)
    vmin = np.select([c1, c3, c2, c0, c4], [b, r, r, b, g], g)
    cat = np.select([c1, c3, c2, c0, c4], [1, 3, 2, 0, 4], 5).astype(np.int32)
    with np.errstate(divide="ignore", invalid="ignore"):
        ratio = ((vmid - vmin).astype(F) / (vmax - vmin).astype(F)).astype(F)
    inc = (cat & 1) == 0
    h = (cat.astype(F) + np.where(inc, ratio, (F(1) - ratio).astype(F))).astype(F)
    h = np.where(vmax == vmin, F(0), h).astype(F)
    h = (h + (F(delta) * F(6)).astype(F)).astype(F)
    for _ in range(3):
        h = np.where(h < 0, (h + F(6)).astype(F), h)
        h = np.where(h >= 6, (h - F(6)).astype(F), h)
    cat2 = h.astype(np.int32)
    ratio2 = (h - cat2.astype(F)).astype(F)
    ratio2 = np.where((cat2 & 1) == 0, ratio2, (F(1) - ratio2).astype(F)).astype(F)
    mid = (vmin + (ratio2 * (vmax - vmin).astype(F)).astype(F)).astype(F)
    sel = [cat2 == 0, cat2 == 1, cat2 == 2, cat2 == 3, cat2 == 4]
    return (np.select(sel, [vmax, mid, vmin, vmin, mid], vmax), np.select(sel, [mid, vmax, vmax, mid, vmin], vmin),
            np.select(sel, [vmin, vmin, mid, vmax, vmax], mid))


def color_augment(image: np.ndarray, brightness_delta: float, contrast_factor: float, saturation_factor: float,
                  hue_delta: float) -> np.ndarray:
    """dataset_builder.py:190-204 with its four random draws given."""
    x = (image.astype(F) + F(brightness_delta)).astype(F)
    mean = (x.astype(np.float64).sum(axis=(0, 1)) / (x.shape[0] * x.shape[1])).astype(F)
    x = (((x - mean).astype(F) * F(contrast_factor)).astype(F) + mean).astype(F)
    h, s, v = _rgb_to_hsv(x[..., 0], x[..., 1], x[..., 2])
    s = np.minimum(F(1), np.maximum(F(0), (s * F(saturation_factor)).astype(F)))
    r, g, b = _hsv_to_rgb(h, s, v)
    r, g, b = _adjust_hue(r, g, b, hue_delta)
    x = np.stack([r, g, b], -1).astype(F)
    lo, hi = x.min(), x.max()
    return ((x - lo).astype(F) / (hi - lo)).astype(F)
