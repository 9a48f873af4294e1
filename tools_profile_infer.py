#!/usr/bin/env python
"""Per-op CUDA-event breakdown of the INFERENCE forward pass (BASELINE config 4: 8-stack, batch 128 per GPU), same
machinery as tools_profile_step.py (hgb_model_profile_all: one event pair around every op of the in-order replay).
    python tools_profile_infer.py --batch 128 --stacks 8 [--out profiles/xxx.md] [--debug k=v,...]"""
import argparse
import json
import os

import torch

import hgb200
from hgb200 import _lib, profiling

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=128)
ap.add_argument("--stacks", type=int, default=8)
ap.add_argument("--out", default=None)
ap.add_argument("--debug", default="")
a = ap.parse_args()
lib = _lib.lib
for kv in [x for x in a.debug.split(",") if x]:
    k, v = kv.split("=")
    lib.hgb_debug_set(int(k), int(v))
model = hgb200.HourglassModel(17, a.stacks, 256, (256, 256, 3), "sigmoid", seed=1)
B = a.batch
img = torch.rand((B, 256, 256, 3), device="cuda")
plan = model._plan(B, False)
for _ in range(3):
    model.forward_device(img, training=False, plan=plan)
torch.cuda.synchronize()
ts = []
for _ in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    model.forward_device(img, training=False, plan=plan)
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
plain_ms = min(ts)
lib.hgb_model_profile_all(plan.handle, 1)
model.forward_device(img, training=False, plan=plan)
torch.cuda.synchronize()
lib.hgb_model_profile_all(plan.handle, 0)
try:
    peaks = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "MEASURED_PEAKS.json")))
except Exception:
    peaks = {}
peak_tf, peak_bw = float(peaks.get("bf16_tflops_sustained", 1400.0)), float(peaks.get("hbm_gbs", 6500.0))
agg, tot = profiling.summarize(plan.handle, 17)
n = sum(r["launches"] for r in agg.values())
rows = profiling.class_table(agg, tot, peak_tf, peak_bw)
lines = [f"# Per-op CUDA-event breakdown of one INFERENCE forward pass: {a.stacks}-stack, batch {B}", "",
         f"Un-profiled forward: {plain_ms:.2f} ms ({B / plain_ms * 1e3:.0f} img/s); sum of per-op event intervals: {tot:.2f} ms over {n} ops "
         f"(in-order replay; roofline fractions against {peak_bw:.0f} GB/s / {peak_tf:.0f} TFLOP/s measured).", "",
         "| op class | launches | total ms | share | avg us | TFLOP/s | GB/s | bound | frac |", "|---|---:|---:|---:|---:|---:|---:|---|---:|"]
for r in rows:
    lines.append(f"| {r['op']} | {r['launches']} | {r['ms']:.3f} | {100 * r['share']:.1f}% | {r['avg_us']:.1f} | {r['tflops']:.0f} | "
                 f"{r['gbps']:.0f} | {r['bound']} | {r['frac']:.2f} |")
text = "\n".join(lines)
print(text)
if a.out:
    open(a.out, "w").write(text + "\n")
