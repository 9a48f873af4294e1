#!/usr/bin/env python
"""End-to-end training throughput FROM TFRECORDS: records -> parse -> nvJPEG -> resize ->
flip/affine/colour augmentation -> target rendering -> 8-stack training step, with and without the background prefetcher
(dataset_builder.py:46 `.prefetch`), against the same step fed from a resident batch.
    python tools_pipeline_bench.py [--batch 128] [--steps 6] [--stacks 8]"""
import argparse
import os
import sys
import tempfile
import time
import types

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=128)
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--stacks", type=int, default=8)
    a = ap.parse_args()
    import cv2
    import torch
    import hgb200
    from hgb200 import tfrecord
    rng = np.random.default_rng(0)
    root = tempfile.mkdtemp(prefix="hgb_tfrec_")
    os.makedirs(os.path.join(root, "train"))
    os.makedirs(os.path.join(root, "valid"))
    n = a.batch * 2
    base = cv2.resize(rng.random((40, 40, 3)).astype(np.float32), (320, 320), interpolation=cv2.INTER_CUBIC)
    payloads = []
    for i in range(n):
        img = np.clip(np.roll(base, i * 7, axis=1) * 255, 0, 255).astype(np.uint8)
        enc = cv2.imencode(".jpg", img, [cv2.IMWRITE_JPEG_QUALITY, 95])[1].tobytes()
        vis = rng.integers(0, 3, 17)
        payloads.append(tfrecord.build_example({
            "ann_id": i, "image_id": i, "image": enc, "image_path": "x.jpg", "coco_url": "u", "width": 320, "height": 320,
            "keypoints/x": (rng.random(17) * 320).astype(np.float32), "keypoints/y": (rng.random(17) * 320).astype(np.float32),
            "keypoints/vis": vis, "keypoints/num": int((vis > 0).sum()), "bbox_x": np.float32(0), "bbox_y": np.float32(0),
            "original_bbox": np.array([0, 0, 320, 320], np.float32)}))
    tfrecord.write_records(os.path.join(root, "train", f"file_train_00-{n}.tfrec"), payloads)
    cfg = types.SimpleNamespace(**{k: getattr(hgb200.default_config, k) for k in dir(hgb200.default_config) if k.isupper()})
    cfg.TRAIN_TFRECORDS_DIR, cfg.VALID_TFRECORDS_DIR = os.path.join(root, "train"), os.path.join(root, "valid")
    cfg.BATCH_SIZE, cfg.SHUFFLE_BUFFER = a.batch, n
    model = hgb200.create_hourglass_model(17, a.stacks, 256, (256, 256, 3), "sigmoid")
    model.compile(optimizer=hgb200.Adam(1e-3), loss=hgb200.loss.weighted_mse)

    def run(next_batch):
        for _ in range(2):
            model.train_on_batch(*next_batch())
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(a.steps):
            model.train_on_batch(*next_batch())
        torch.cuda.synchronize()
        return a.batch * a.steps / (time.perf_counter() - t0)

    def run_fit_loop(next_batch):
        """the loop HourglassModel.fit runs: step i's losses are read after step i+1 has been enqueued"""
        for _ in range(2):
            model.train_on_batch(*next_batch())
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        pending = None
        for _ in range(a.steps):
            nxt = model.train_on_batch_deferred(*next_batch())
            if pending is not None:
                pending.result()
            pending = nxt
        pending.result()
        torch.cuda.synchronize()
        return a.batch * a.steps / (time.perf_counter() - t0)

    out = {}
    ds0 = hgb200.dataset_builder.DatasetBuilder(cfg, seed=0, prefetch=0).build_datasets()[0]
    resident = next(ds0)
    out["resident batch (step only)"] = run(lambda: resident)
    out["from TFRecords, prefetch=0"] = run(lambda: next(ds0))
    ds2 = hgb200.dataset_builder.DatasetBuilder(cfg, seed=0, prefetch=2).build_datasets()[0]
    out["from TFRecords, prefetch=2"] = run(lambda: next(ds2))
    out["resident batch, fit loop (deferred loss reads)"] = run_fit_loop(lambda: resident)
    out["from TFRecords, prefetch=2, fit loop"] = run_fit_loop(lambda: next(ds2))
    ds2.close()
    for k, v in out.items():
        print(f"{k:50s} {v:9.1f} img/s   ({a.stacks}-stack, batch {a.batch}, {a.steps} timed steps)")


if __name__ == "__main__":
    main()
