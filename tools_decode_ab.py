#!/usr/bin/env python
"""A/B of the decode kernel's launch geometry (hgb_debug_set 24 = CTAs per sample, 28 = 10*stages + vectors per thread and
chunk) at one shape, CUDA events, clean-L2 flush between launches.   python tools_decode_ab.py [--batch 1024] [--hw 64]"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
import hgb200  # noqa: F401
from hgb200 import _lib
from hgb200._lib import lib, check, ptr, stream_ptr

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=1024)
ap.add_argument("--hw", type=int, default=64)
ap.add_argument("--profile", type=int, default=0, help="1: bracket one default-geometry f32 launch with cudaProfilerStart/Stop")
a = ap.parse_args()
B, H, K = a.batch, a.hw, 17
g = torch.Generator(device="cuda").manual_seed(1)
yp = torch.rand((B, H, H, K), device="cuda", generator=g)
ypb = yp.to(torch.bfloat16)
idx = torch.empty((B, K, 4), dtype=torch.int32, device="cuda")
kp = torch.empty((B, K, 3), dtype=torch.float32, device="cuda")
flush = torch.empty(64 << 20, dtype=torch.float32, device="cuda")
flush2 = torch.zeros(64 << 20, dtype=torch.float32, device="cuda")
sink = torch.zeros((), device="cuda")
st = stream_ptr()


def timed(fn, iters=5):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(iters):
        flush.fill_(1.0); sink.add_(flush2.sum())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best * 1e-3


if a.profile:
    flush.fill_(1.0); sink.add_(flush2.sum()); torch.cuda.synchronize()
    torch.cuda.profiler.start()
    check(lib.hgb_decode(ptr(yp), _lib.F32, B, H, H, K, 1e-6, 2, ptr(idx), ptr(kp), st))
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    raise SystemExit(0)
print(f"decode v2, batch {B}, {H}x{H}x{K}: GB/s (us) per geometry")
for split in (0, 1, 2, 4, 8):
    for geo in (0, 44, 64, 28, 48, 38):
        lib.hgb_debug_set(24, split); lib.hgb_debug_set(28, geo)
        try:
            tf = timed(lambda: check(lib.hgb_decode(ptr(yp), _lib.F32, B, H, H, K, 1e-6, 2, ptr(idx), ptr(kp), st)))
            tb = timed(lambda: check(lib.hgb_decode(ptr(ypb), _lib.BF16, B, H, H, K, 1e-6, 2, ptr(idx), ptr(kp), st)))
            print(f"split {split} (0 = auto)  stages*10+iters {geo:2d} (0 = 44): f32 {yp.numel() * 4 / tf / 1e9:6.0f} ({tf * 1e6:6.1f})   "
                  f"bf16 {yp.numel() * 2 / tb / 1e9:6.0f} ({tb * 1e6:6.1f})")
        except Exception as ex:
            print(f"split {split} geo {geo}: {ex}")
lib.hgb_debug_set(24, 0); lib.hgb_debug_set(28, 0)
