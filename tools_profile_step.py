#!/usr/bin/env python
"""Per-op time breakdown of one training step using the library's CUDA-event profiler
(hgb_model_profile_all): cheap enough for the full batch-256 step, where an ncu launch list is not.
    python tools_profile_step.py --batch 256 --stacks 8 [--out profiles/xxx.md]"""
import argparse
import collections
import ctypes as C
import sys

import torch

import hgb200
from hgb200 import _lib, ops

NAMES = ["F_IM2COL", "F_CONV", "F_BN", "F_POOL", "F_UPADD", "F_HEAD", "B_BN_REDUCE", "B_BN_APPLY", "B_WGRAD", "B_DGRAD",
         "B_RELU_MASK", "B_COLSUM", "B_POOL", "B_UPADD", "B_HEAD"]
ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--stacks", type=int, default=8)
ap.add_argument("--out", default=None)
ap.add_argument("--debug", default="", help="comma-separated key=value pairs for hgb_debug_set")
a = ap.parse_args()
lib = _lib.lib
for kv in [x for x in a.debug.split(",") if x]:
    k, v = kv.split("=")
    lib.hgb_debug_set(int(k), int(v))
model = hgb200.HourglassModel(17, a.stacks, 256, (256, 256, 3), "sigmoid", seed=1)
model.compile(optimizer=hgb200.Adam(1e-3), loss=hgb200.loss.weighted_mse)
B = a.batch
img = torch.rand((B, 256, 256, 3), device="cuda")
tg = ops.render_targets(torch.rand((B, 17), device="cuda") * 64, torch.rand((B, 17), device="cuda") * 64,
                        torch.randint(0, 3, (B, 17), device="cuda", dtype=torch.int32), 64, 64)
plan = model._plan(B, True)
for _ in range(3):
    model.train_step_device(img, tg)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
model.train_step_device(img, tg)
e1.record()
torch.cuda.synchronize()
plain_ms = e0.elapsed_time(e1)
lib.hgb_model_profile_all(plan.handle, 1)
model.train_step_device(img, tg)
torch.cuda.synchronize()
lib.hgb_model_profile_all(plan.handle, 0)
from hgb200 import profiling
import json, os
try:
    peaks = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "MEASURED_PEAKS.json")))
except Exception:
    peaks = {}
peak_tf, peak_bw = float(peaks.get("bf16_tflops_sustained", 1400.0)), float(peaks.get("hbm_gbs", 6500.0))
agg_d, tot = profiling.summarize(plan.handle, 17)
n = sum(r["launches"] for r in agg_d.values())
rows = profiling.class_table(agg_d, tot, peak_tf, peak_bw)
agg = {r["op"]: [r["launches"], r["ms"], 0, 0] for r in rows}
lines = [f"# Per-op CUDA-event breakdown of one training step: {a.stacks}-stack, batch {B}",
         "", f"Un-profiled step: {plain_ms:.2f} ms; sum of per-op event intervals: {tot:.2f} ms over {n} ops "
         "(event pairs on the launching stream around every op of the in-order replay; bytes = every operand tensor of the op "
         f"moved once, hgb200/profiling.py; roofline fractions against {peak_bw:.0f} GB/s / {peak_tf:.0f} TFLOP/s measured).", "",
         "| op class | launches | total ms | share | avg us | TFLOP/s | GB/s | bound | frac |", "|---|---:|---:|---:|---:|---:|---:|---|---:|"]
for r in rows:
    lines.append(f"| {r['op']} | {r['launches']} | {r['ms']:.3f} | {100 * r['share']:.1f}% | {r['avg_us']:.1f} | {r['tflops']:.0f} | "
                 f"{r['gbps']:.0f} | {r['bound']} | {r['frac']:.2f} |")
by_type = collections.defaultdict(float)
for k, v in agg.items():
    by_type[k.split()[0]] += v[1]
lines += ["", "| op type | total ms | share |", "|---|---:|---:|"]
for k, v in sorted(by_type.items(), key=lambda kv: -kv[1]):
    lines.append(f"| {k} | {v:.3f} | {100 * v / tot:.1f}% |")
text = "\n".join(lines)
print(text)
if a.out:
    open(a.out, "w").write(text + "\n")
