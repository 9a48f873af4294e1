#!/usr/bin/env python
"""Input-path kernels at BASELINE sizes: batch 256 of 256x256x3 float32 images, timed with
CUDA events through the C ABI on preallocated buffers, L2 flushed between launches.  Reported as achieved HBM GB/s of the
bytes each op moves, next to the reference's CPU path for the same op on this box's host cores (single thread, which
is how the reference runs them: one Python callback per example, dataset_builder.py:187,237):

  crop_resize    uint8 640x480 frames -> (256,256,3) f32 : read 3 B/px of the source + write 12 B/px          cv2.resize
  augment_affine flip + affine warp                       : read + write 12 B/px                               cv2.warpAffine
  color_augment  brightness/contrast/saturation/hue/norm  : 3 reads + 2 writes of 12 B/px (three passes)       numpy restatement (bench.py only)
  jpeg_decode    nvJPEG, 256x256 4:2:0 q95                : decoded MB/s and images/s                          cv2.imdecode
  train_label    augment_1 + augment_2 + target rendering (make_train_label_batch), images/s

    python tools_input_bench.py [--json] [--batch 256]"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def _cpu_time(fn, min_seconds=1.0, max_iters=200):
    fn()
    t0, n = time.perf_counter(), 0
    while n < max_iters:
        fn()
        n += 1
        if time.perf_counter() - t0 > min_seconds:
            break
    return (time.perf_counter() - t0) / n


def sweep(batch=256, iters=5, cpu=True, cpu_color_fn=None):
    import cv2
    import torch
    import hgb200  # noqa: F401  (registers the package alias)
    from hgb200 import _lib, dataset_builder as db, tfrecord
    from hgb200._lib import check, lib, ptr, stream_ptr

    cv2.setNumThreads(1)
    rng = np.random.default_rng(0)
    N, H, W = batch, 256, 256
    flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device="cuda")
    st = stream_ptr()

    def timed(fn):
        fn()
        torch.cuda.synchronize()
        tot = 0.0
        for _ in range(iters):
            flush.fill_(1.0)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            tot += e0.elapsed_time(e1)
        return tot / iters * 1e-3

    rows = {}
    img_bytes = H * W * 3 * 4
    gen = torch.Generator(device="cuda").manual_seed(0)

    # ---- crop_resize: N distinct 640x480 uint8 frames, whole image -> 256x256
    sh, sw = 480, 640
    frames = torch.randint(0, 256, (N, sh, sw, 3), device="cuda", generator=gen, dtype=torch.uint8)
    table = torch.tensor([frames[i].data_ptr() for i in range(N)], dtype=torch.int64).cuda()
    hw = torch.tensor([[sh, sw]] * N, dtype=torch.int32).cuda()
    out = torch.empty((N, H, W, 3), dtype=torch.float32, device="cuda")
    t = timed(lambda: check(lib.hgb_crop_resize(ptr(table), ptr(hw), _lib.U8, None, N, H, W, ptr(out), st)))
    rows["crop_resize u8 640x480"] = {"s": t, "bytes": N * (sh * sw * 3 + img_bytes)}
    # ---- augment_affine
    images = torch.rand((N, H, W, 3), device="cuda", generator=gen)
    d = db.draw_augmentation(rng, N)
    inv = np.stack([db._opencv_inverse(db.affine_matrix(H, W, float(s), float(r), 0.5)) for s, r in zip(d["scale"], d["rotate_deg"])])
    inv_t = torch.as_tensor(inv.reshape(N, 6)).cuda()
    flip_t = torch.as_tensor(d["flip"].astype(np.int32)).cuda()
    t = timed(lambda: check(lib.hgb_augment_affine(ptr(images), ptr(inv_t), ptr(flip_t), N, H, W, ptr(out), st)))
    rows["augment_affine"] = {"s": t, "bytes": 2 * N * img_bytes}
    # ---- color_augment (in place on `out`)
    params = torch.as_tensor(np.stack([d["brightness_delta"], d["contrast_factor"], d["saturation_factor"], d["hue_delta"]], 1).astype(np.float32)).cuda()
    ws = torch.empty(int(lib.hgb_color_workspace_bytes(N)), dtype=torch.uint8, device="cuda")
    work = images.clone()

    def color():
        check(lib.hgb_color_augment(ptr(work), ptr(params), N, H, W, ptr(ws), st))
    t = timed(color)
    rows["color_augment"] = {"s": t, "bytes": 5 * N * img_bytes}
    # ---- whole training label
    kx = torch.rand((N, 17), device="cuda", generator=gen) * 72 - 4
    ky = torch.rand((N, 17), device="cuda", generator=gen) * 72 - 4
    kv = torch.randint(0, 3, (N, 17), device="cuda", generator=gen, dtype=torch.int32)
    t = timed(lambda: db.make_train_label_batch(images, kx, ky, kv, d))
    rows["train_label (aug1+aug2+render)"] = {"s": t, "images_per_s": N / t}
    # ---- nvJPEG
    small = cv2.resize(rng.random((34, 34, 3)).astype(np.float32), (W, H), interpolation=cv2.INTER_CUBIC)
    photo = np.clip(small * 255, 0, 255).astype(np.uint8)
    enc = cv2.imencode(".jpg", photo, [cv2.IMWRITE_JPEG_QUALITY, 95])[1].tobytes()
    n_jpeg = min(N, 64)
    streams = [enc] * n_jpeg
    tfrecord.decode_jpeg_batch(streams)
    t0 = time.perf_counter()
    for _ in range(3):
        tfrecord.decode_jpeg_batch(streams)
    tj = (time.perf_counter() - t0) / 3
    rows["jpeg_decode nvJPEG 256x256"] = {"s": tj, "images_per_s": n_jpeg / tj, "bytes": n_jpeg * H * W * 3, "host_timed": True}

    if cpu:
        f = frames[0].cpu().numpy()
        rows["crop_resize u8 640x480"]["cpu_s_per_image"] = _cpu_time(lambda: cv2.resize(f.astype(np.float32) * np.float32(1 / 255), (W, H), interpolation=cv2.INTER_LINEAR))
        im = images[0].cpu().numpy()
        m = db.affine_matrix(H, W, 1.1, 17.0, 0.5)[:2]
        rows["augment_affine"]["cpu_s_per_image"] = _cpu_time(lambda: cv2.warpAffine(im, m, dsize=(W, H), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_CONSTANT, borderValue=0))
        if cpu_color_fn is not None:      # the numpy restatement lives under oracle/: only bench.py's CPU leg hands it in
            rows["color_augment"]["cpu_s_per_image"] = _cpu_time(lambda: cpu_color_fn(im, 0.1, 1.3, 1.1, 0.05), min_seconds=2.0, max_iters=20)
        arr = np.frombuffer(enc, np.uint8)
        rows["jpeg_decode nvJPEG 256x256"]["cpu_s_per_image"] = _cpu_time(lambda: cv2.imdecode(arr, cv2.IMREAD_COLOR))
    for k, r in rows.items():
        r["batch"] = n_jpeg if k.startswith("jpeg") else N
        if "bytes" in r:
            r["gbps"] = r["bytes"] / r["s"] / 1e9
        if "cpu_s_per_image" in r:
            r["cpu_images_per_s"] = 1.0 / r["cpu_s_per_image"]
            r["gpu_images_per_s"] = r["batch"] / r["s"]
    return rows


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--json", action="store_true")
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU reference timings (profiling runs)")
    a = ap.parse_args()
    rows = sweep(a.batch, iters=a.iters, cpu=not a.no_cpu)
    if a.json:
        print(json.dumps(rows))
    else:
        print(f"{'op':36s} {'batch':>5s} {'us':>9s} {'GB/s':>8s} {'GPU img/s':>11s} {'CPU img/s (1 thread)':>21s}")
        for k, r in rows.items():
            print(f"{k:36s} {r['batch']:5d} {r['s'] * 1e6:9.1f} {r.get('gbps', float('nan')):8.1f} "
                  f"{r.get('gpu_images_per_s', r.get('images_per_s', float('nan'))):11.0f} {r.get('cpu_images_per_s', float('nan')):21.1f}")
