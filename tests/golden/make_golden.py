"""Generate tests/golden/*.npz by running the REFERENCE's own numpy functions.

Run in the build container only (needs /root/reference, which the GPU box lacks):
    python tests/golden/make_golden.py
TensorFlow / matplotlib / pycocotools / imgaug are absent here, so stub modules are
planted in sys.modules first; only numpy code paths execute.
The fixtures are committed; tests never read /root/reference.
"""
import contextlib
import io
import os
import sys
import types

import numpy as np

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def import_reference():
    def stub(name, **attrs):
        m = types.ModuleType(name)
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[name] = m
        return m

    tf = stub("tensorflow")
    tf.data = types.SimpleNamespace(Dataset=object)
    stub("matplotlib")
    stub("matplotlib.pyplot")
    stub("matplotlib.patches")
    stub("pycocotools")
    stub("pycocotools.coco", COCO=object)
    stub("pycocotools.cocoeval", COCOeval=object)
    stub("imgaug")
    stub("imgaug.augmenters")
    stub("imgaug.augmentables", Keypoint=object, KeypointsOnImage=object)
    sys.path.insert(0, REF)
    import dataset_builder  # noqa
    import eval as ref_eval  # noqa
    from utilities import data_utils  # noqa
    return data_utils, dataset_builder, ref_eval


def decode_cases(rng):
    """(N,64,64,17) float32 maps with planted peaks: interior, borders, corners, ties,
    all-negative neighbourhoods, sub-threshold, all-zero, NaN."""
    H = W = 64
    K = 17
    maps = []
    # 0: uniform random
    maps.append(rng.random((H, W, K), dtype=np.float32))
    # 1: planted positions per joint
    m = rng.random((H, W, K), dtype=np.float32) * 0.5
    spots = [(0, 0), (0, 63), (63, 0), (63, 63), (0, 5), (5, 0), (63, 17), (17, 63),
             (1, 1), (62, 62), (30, 31), (1, 0), (0, 1), (20, 20), (40, 7), (7, 40), (33, 33)]
    for k, (y, x) in enumerate(spots):
        m[y, x, k] = 2.0 + k
    maps.append(m)
    # 2: exact ties (lowest flat index wins) + second max in different patch cells
    m = np.zeros((H, W, K), dtype=np.float32)
    for k in range(K):
        y, x = 3 + 3 * k, 60 - 3 * k
        m[y, x, k] = 1.0
        m[min(y + 5, 63), max(x - 7, 0), k] = 1.0           # tie later in scan order
        dy, dx = divmod(k % 9, 3)
        if (dy, dx) != (1, 1):
            m[y - 1 + dy, x - 1 + dx, k] = 0.5                # 2nd max location in the 3x3
    maps.append(m)
    # 3: all-negative neighbourhood around an interior peak, and negative maps
    m = -rng.random((H, W, K), dtype=np.float32) - 0.1
    for k in range(K):
        if k % 2 == 0:
            m[10 + k, 12 + k, k] = 0.75
    maps.append(m)
    # 4: all zero
    maps.append(np.zeros((H, W, K), dtype=np.float32))
    # 5: sub-threshold peaks (1e-7 < 1e-6) and tiny negatives
    m = np.zeros((H, W, K), dtype=np.float32)
    for k in range(K):
        m[5 + k, 9, k] = 1e-7 if k % 2 else 2e-6
        m[5 + k, 10, k] = 5e-8
    maps.append(m)
    # 6: NaN present (np.argmax treats NaN as the maximum)
    m = rng.random((H, W, K), dtype=np.float32)
    m[13, 14, 3] = np.nan
    m[0, 0, 5] = np.nan
    m[40, 41, 5] = np.nan
    maps.append(m)
    # 7: border peaks with large neighbours in every direction (patch clipping quirks)
    m = rng.random((H, W, K), dtype=np.float32) * 0.1
    border = [(0, 10), (10, 0), (63, 10), (10, 63), (0, 0), (0, 63), (63, 0), (63, 63)]
    for k in range(K):
        y, x = border[k % 8]
        m[y, x, k] = 5.0
        for dy in (-1, 0, 1):
            for dx in (-1, 0, 1):
                yy, xx = y + dy, x + dx
                if (dy or dx) and 0 <= yy < H and 0 <= xx < W:
                    m[yy, xx, k] = 1.0 + 0.1 * ((dy + 1) * 3 + dx + 1) * (1 if k < 8 else -1) + (0 if k < 8 else 1.0)
    maps.append(m)
    # 8..9: sigmoid-like smooth maps
    for _ in range(2):
        z = rng.standard_normal((H, W, K)).astype(np.float32)
        maps.append((1.0 / (1.0 + np.exp(-z))).astype(np.float32))
    return np.stack(maps)


def main():
    du, dsb, ref_eval = import_reference()
    rng = np.random.default_rng(1234)

    # ---------------- decode ----------------
    hm = decode_cases(rng)
    N, H, W, K = hm.shape
    thr_list = [1e-6, 0.1]
    out = {"heatmaps": hm}
    for ti, thr in enumerate(thr_list):
        v1 = np.zeros((N, K, 3), np.float32)
        v2 = np.zeros((N, K, 3), np.float32)
        mut = np.zeros_like(hm)
        for n in range(N):
            v1[n] = du.heatmaps_to_keypoints_v1(hm[n].copy(), thr)
            work = hm[n].copy()
            v2[n] = du.heatmaps_to_keypoints_v2(work, thr)
            mut[n] = work
        out[f"v1_thr{ti}"] = v1
        out[f"v2_thr{ti}"] = v2
        if ti == 0:                          # caller's-array mutation, stored sparsely
            d = np.nonzero(~((mut == hm) | (np.isnan(mut) & np.isnan(hm))))
            out["v2_mut_idx"] = np.stack(d).astype(np.int32)
            out["v2_mut_val"] = mut[d]
    out["thresholds"] = np.array(thr_list, np.float64)
    np.savez_compressed(os.path.join(OUT, "decode_golden.npz"), **out)

    # ---------------- target rendering ----------------
    builder = dsb.DatasetBuilder.__new__(dsb.DatasetBuilder)
    builder.label_shape = (64, 64, 17)
    builder.num_keypoints = 17
    B = 12
    kx = rng.uniform(-4, 68, size=(B, 17)).astype(np.float32)
    ky = rng.uniform(-4, 68, size=(B, 17)).astype(np.float32)
    kv = rng.choice([0, 1, 2], p=[.2, .3, .5], size=(B, 17)).astype(np.int64)
    # planted edge cases: borders, clipping, strict inequalities, v=0
    kx[0, :8] = [10.7, 1.9, 0.5, 63.9, 0.999, 1.0, 63.0, 64.0]
    ky[0, :8] = [20.2, 62.5, 10.0, 63.9, 30.0, 1.0, 1.0, 5.0]
    kv[0, :8] = [2, 1, 2, 2, 2, 1, 1, 2]
    kx[1, :4] = [2.5, 61.2, 3.0, -0.5]
    ky[1, :4] = [2.5, 61.9, 60.0, 12.0]
    kv[1, :4] = [1, 2, 0, 2]
    targets = np.stack([builder.np_gen_heatmaps(kx[b], ky[b], kv[b]) for b in range(B)])
    g = np.zeros((7, 7), np.float32)
    img = np.zeros((7, 7), np.float32)
    g[:] = du.gaussian(img, (3, 3))
    # 128x128 labels (config 5 sweep)
    builder2 = dsb.DatasetBuilder.__new__(dsb.DatasetBuilder)
    builder2.label_shape = (128, 128, 17)
    builder2.num_keypoints = 17
    kx2 = (kx[:3] * 2).astype(np.float32)
    ky2 = (ky[:3] * 2).astype(np.float32)
    targets128 = np.stack([builder2.np_gen_heatmaps(kx2[b], ky2[b], kv[b]) for b in range(3)])
    # store the 128 maps as sparse (indices + values) to keep the fixture small
    nz = np.nonzero(targets128)
    np.savez_compressed(os.path.join(OUT, "render_golden.npz"),
                        kps_x=kx, kps_y=ky, kps_v=kv, targets=targets, gaussian7=g,
                        kps_x128=kx2, kps_y128=ky2,
                        t128_idx=np.stack(nz).astype(np.int32), t128_val=targets128[nz])

    # ---------------- PCK + bbox helpers ----------------
    labels = ["j%d" % i for i in range(17)]
    P = 40
    preds = []
    for n in range(P):
        bbox = rng.uniform(20, 300, size=4)
        xs_gt = rng.uniform(0, 400, size=17)
        ys_gt = rng.uniform(0, 400, size=17)
        diam = np.sqrt(bbox[2] ** 2 + bbox[3] ** 2)
        r = rng.uniform(0, 0.1, size=17) * diam
        th = rng.uniform(0, 2 * np.pi, size=17)
        xs_pred = xs_gt + r * np.cos(th)
        ys_pred = ys_gt + r * np.sin(th)
        vs = rng.choice([0, 1, 2], p=[.2, .3, .5], size=17)
        if n == 0:
            vs[:] = 2                        # every joint visible at least once
        if n == 1:                           # exact-boundary cases: dist == threshold
            bbox[2], bbox[3] = 60.0, 80.0    # diameter 100 -> threshold 5
            xs_pred[:3] = xs_gt[:3] + np.array([3.0, 5.0, 0.0])
            ys_pred[:3] = ys_gt[:3] + np.array([4.0, 0.0, 5.0])
            vs[:3] = 2
        preds.append({"original_bbox": bbox.tolist(), "xs/pred": xs_pred.tolist(), "ys/pred": ys_pred.tolist(),
                      "xs/gt": xs_gt.tolist(), "ys/gt": ys_gt.tolist(), "vs": [int(v) for v in vs]})
    with contextlib.redirect_stdout(io.StringIO()):
        stats005 = ref_eval.eval_PCK(preds, labels, 0.05)
        stats002 = ref_eval.eval_PCK(preds, labels, 0.02)
    sq1 = du.transform_bbox_square([603.15, 125.6, 36.85, 66.16])
    sq2 = du.transform_bbox_square([163.73, 126.42, 265.69, 480.4], 1.25)
    nx = rng.random(17)
    ny = rng.random(17)
    ux, uy = ref_eval._undo_bbox(12.5, -3.25, 200, 180, nx, ny)
    np.savez_compressed(os.path.join(OUT, "score_golden.npz"),
                        bbox=np.array([p["original_bbox"] for p in preds]),
                        xs_pred=np.array([p["xs/pred"] for p in preds]), ys_pred=np.array([p["ys/pred"] for p in preds]),
                        xs_gt=np.array([p["xs/gt"] for p in preds]), ys_gt=np.array([p["ys/gt"] for p in preds]),
                        vs=np.array([p["vs"] for p in preds]),
                        pck005=np.array(stats005), pck002=np.array(stats002),
                        square1=np.array(sq1), square2=np.array(sq2),
                        undo_in=np.stack([nx, ny]), undo_out=np.stack([ux, uy]))
    print("wrote fixtures to", OUT)
    for f in sorted(os.listdir(OUT)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
