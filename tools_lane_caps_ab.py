#!/usr/bin/env python
"""Grid search of the two side-lane CTA caps inside the real multi-lane training step:
  key 9  = SMs the skip lanes' persistent GEMMs leave free (default 20 -> 128 CTAs)
  key 20 = CTAs of a weight gradient on its side lane (default 64)
    python tools_lane_caps_ab.py --batch 32 --reserves 20,40,60,80 --wgrads 24,32,48,64"""
import argparse, sys
import torch
sys.path.insert(0, '.')
import hgb200
from hgb200 import ops, _lib
ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--reserves", default="20,40,60,80")
ap.add_argument("--wgrads", default="24,32,48,64")
ap.add_argument("--steps", type=int, default=8)
ap.add_argument("--wgrad-lanes", type=int, default=0, help="key 35 (set before the plan is built): streams the main chain's weight gradients are dealt over")
a = ap.parse_args()
lib = _lib.lib
lib.hgb_debug_set(35, a.wgrad_lanes)
B = a.batch
model = hgb200.HourglassModel(17, 8, 256, (256, 256, 3), "sigmoid", seed=1)
model.compile(optimizer=hgb200.Adam(1e-3), loss=hgb200.loss.weighted_mse)
img = torch.rand((B, 256, 256, 3), device="cuda")
tg = ops.render_targets(torch.rand((B, 17), device="cuda") * 64, torch.rand((B, 17), device="cuda") * 64,
                        torch.randint(0, 3, (B, 17), device="cuda", dtype=torch.int32), 64, 64)
for _ in range(3):
    model.train_step_device(img, tg)


def timed():
    model.train_step_device(img, tg)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        model.train_step_device(img, tg)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / a.steps


res = {}
for rep in range(2):
    for r in [int(x) for x in a.reserves.split(",")]:
        for w in [int(x) for x in a.wgrads.split(",")]:
            lib.hgb_debug_set(9, r)
            lib.hgb_debug_set(20, w)
            res.setdefault((r, w), []).append(timed())
print(f"batch {B}, weight-gradient lanes {a.wgrad_lanes or 'default'}: ms/step (best of 2 x {a.steps} steps); rows = skip-lane SM reserve, columns = weight-gradient CTA cap")
ws = [int(x) for x in a.wgrads.split(",")]
print("reserve " + " ".join(f"{w:>8d}" for w in ws))
for r in [int(x) for x in a.reserves.split(",")]:
    print(f"{r:7d} " + " ".join(f"{min(res[(r, w)]):8.2f}" for w in ws))
