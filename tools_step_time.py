#!/usr/bin/env python
"""Wall-clock-free step timing of the real multi-lane training step (CUDA events) for A/B runs of two builds on the SAME box:
    python tools_step_time.py --batches 256,32            (this tree)
    cd ab_old && python ../tools_step_time.py ...          (an older tree extracted there with its own libhgb200.so)"""
import argparse, os, sys
sys.path.insert(0, os.getcwd())
import torch
import hgb200
from hgb200 import ops, _lib
ap = argparse.ArgumentParser()
ap.add_argument("--batches", default="256,32")
ap.add_argument("--steps", type=int, default=6)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--debug", default="")
a = ap.parse_args()
for kv in [x for x in a.debug.split(",") if x]:
    k, v = kv.split("=")
    _lib.lib.hgb_debug_set(int(k), int(v))
print("library:", _lib.LIB_PATH)
for B in [int(b) for b in a.batches.split(",")]:
    model = hgb200.HourglassModel(17, 8, 256, (256, 256, 3), "sigmoid", seed=1)
    model.compile(optimizer=hgb200.Adam(1e-3), loss=hgb200.loss.weighted_mse)
    img = torch.rand((B, 256, 256, 3), device="cuda")
    tg = ops.render_targets(torch.rand((B, 17), device="cuda") * 64, torch.rand((B, 17), device="cuda") * 64,
                            torch.randint(0, 3, (B, 17), device="cuda", dtype=torch.int32), 64, 64)
    for _ in range(3):
        model.train_step_device(img, tg)
    best = []
    for rep in range(a.reps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.steps):
            model.train_step_device(img, tg)
        e1.record()
        torch.cuda.synchronize()
        best.append(e0.elapsed_time(e1) / a.steps)
    print(f"batch {B}: ms/step {' '.join(f'{t:.2f}' for t in best)}  (min {min(best):.2f})", flush=True)
    del model, img, tg
    torch.cuda.empty_cache()
