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
n = lib.hgb_model_profile_count(plan.handle)
agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])   # count, ms, flops, bytes
info, ms = (C.c_int * 8)(), C.c_double()
cinfo, coffs = (C.c_int * 8)(), (C.c_int64 * 2)()
off, dims = C.c_int64(), (C.c_int * 4)()


def act_dims(i):
    lib.hgb_model_act_info(plan.handle, i, C.byref(off), C.byref(dims))
    return tuple(dims)


tot = 0.0
for i in range(n):
    _lib.check(lib.hgb_model_profile_op(plan.handle, i, C.byref(info), C.byref(ms)))
    ty, conv, bn, a0, a1, a2, a3, flag = tuple(info)
    key, flops, byt = NAMES[ty], 0.0, 0.0
    if conv >= 0 and ty in (1, 8, 9):
        lib.hgb_model_conv_detail(plan.handle, conv, C.byref(cinfo), C.byref(coffs))
        ks, taps, cin, cout, cinp, coutp = cinfo[0], cinfo[1], cinfo[2], cinfo[3], cinfo[4], cinfo[5]
        nn, hh, ww, _ = act_dims(a0)
        key += f" k{ks} {cin}->{cout} @{hh}"
        flops = 2.0 * nn * hh * ww * taps * cin * cout
        byt = 2.0 * nn * hh * ww * (cinp + coutp)
    elif a0 >= 0:
        nn, hh, ww, cc = act_dims(a0 if ty != 5 else a0)
        key += f" C{cc} @{hh}"
        mult = {2: 2.3, 6: 2, 7: 3, 3: 1.25, 4: 2.25, 10: 3, 11: 1, 12: 2.25, 13: 1.25}.get(ty, 2)
        byt = 2.0 * nn * hh * ww * cc * mult
    agg[key][0] += 1
    agg[key][1] += ms.value
    agg[key][2] += flops
    agg[key][3] += byt
    tot += ms.value
lines = [f"# Per-op CUDA-event breakdown of one training step: {a.stacks}-stack, batch {B}",
         "", f"Un-profiled step: {plain_ms:.2f} ms; sum of per-op event intervals: {tot:.2f} ms over {n} ops "
         "(event pairs on the launching stream around every op; bytes are the op's algorithmic activation traffic).", "",
         "| op class | launches | total ms | share | avg us | TFLOP/s | GB/s |", "|---|---:|---:|---:|---:|---:|---:|"]
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    tf = v[2] / (v[1] * 1e-3) / 1e12 if v[2] else 0
    gb = v[3] / (v[1] * 1e-3) / 1e9 if v[3] else 0
    lines.append(f"| {k} | {v[0]} | {v[1]:.3f} | {100 * v[1] / tot:.1f}% | {1e3 * v[1] / v[0]:.1f} | {tf:.0f} | {gb:.0f} |")
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
