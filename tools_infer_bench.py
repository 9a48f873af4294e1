#!/usr/bin/env python
"""BASELINE.json config 4 on one GPU's shard: 8-stack inference (inference-mode BN) + v2 heat-map decode + PCK and
OKS-similarity scoring, batch 128 (= 1024 / 8 GPUs; the path shards by batch with no collective), device-resident
inputs, CUDA events.  Reports images/s of the whole chain and the time of each stage.
    python tools_infer_bench.py [--batch 128] [--iters 5]"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np
import torch
import hgb200
from hgb200 import ops

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=128)
ap.add_argument("--iters", type=int, default=5)
a = ap.parse_args()
B, K = a.batch, 17
model = hgb200.HourglassModel(K, 8, 256, (256, 256, 3), "sigmoid", seed=1)
g = torch.Generator(device="cuda").manual_seed(0)
images = torch.rand((B, 256, 256, 3), device="cuda", generator=g)
gt_x = torch.rand((B, K), device="cuda", generator=g, dtype=torch.float64) * 200
gt_y = torch.rand((B, K), device="cuda", generator=g, dtype=torch.float64) * 200
vis = torch.randint(0, 3, (B, K), device="cuda", generator=g, dtype=torch.int32)
bbox = torch.rand((B, 4), device="cuda", generator=g, dtype=torch.float64) * 100 + 100
area = bbox[:, 2] * bbox[:, 3]


def chain():
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    ev[0].record()
    heat = model.forward_device(images, training=False)[-1]
    ev[1].record()
    _idx, kp = ops.decode_batch(heat, 1e-6, 2)
    ev[2].record()
    xs = kp[:, :, 0].double() / 64 * bbox[:, 2:3] + bbox[:, 0:1]       # eval.py:114-126 on the device
    ys = kp[:, :, 1].double() / 64 * bbox[:, 3:4] + bbox[:, 1:2]
    correct, visible = ops.pck_counts(xs, ys, gt_x, gt_y, vis, bbox[:, 2:4])
    oks = ops.oks_similarity(xs, ys, gt_x, gt_y, vis, area, bbox)
    ev[3].record()
    torch.cuda.synchronize()
    return [ev[i].elapsed_time(ev[i + 1]) for i in range(3)], float(oks.mean()), int(visible.sum())


for _ in range(3):
    chain()
tot = np.zeros(3)
for _ in range(a.iters):
    t, _o, _v = chain()
    tot += np.array(t)
tot /= a.iters
print(f"8-stack inference + decode v2 + PCK/OKS, batch {B}: forward {tot[0]:.2f} ms, decode {tot[1] * 1e3:.0f} us, scoring (incl. host read-back of "
      f"34 counters) {tot[2] * 1e3:.0f} us -> {B / tot.sum() * 1e3:.0f} img/s per GPU "
      f"(forward alone {B / tot[0] * 1e3:.0f} img/s = {68.87 * B / tot[0]:.0f} TFLOP/s)")
