#!/usr/bin/env python
"""Microbenchmark of the tcgen05 convolution kernels at the hourglass layer shapes (stand-alone C-ABI
entry points hgb_conv_gemm / hgb_conv_wgrad; tensors are far larger than the 126 MB L2).
    python tools_conv_bench.py [--batch 256] [--only k3]"""
import argparse

import torch

import hgb200
from hgb200 import ops

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--iters", type=int, default=10)
ap.add_argument("--only", default="")
a = ap.parse_args()
B = a.batch
CASES = [("k3 128->128 @64", 3, 128, 128, 64), ("k1 128->256 @64", 1, 128, 256, 64), ("k1 256->128 @64", 1, 256, 128, 64),
         ("k1 256->256 @64", 1, 256, 256, 64), ("k3 128->128 @32", 3, 128, 128, 32), ("k3 128->128 @16", 3, 128, 128, 16)]


def timeit(fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / a.iters * 1e-3


print(f"{'case':20s} {'op':12s} {'us':>9s} {'TFLOP/s':>9s} {'GB/s':>8s}")
for name, k, cin, cout, h in CASES:
    if a.only and a.only not in name:
        continue
    x = (torch.randn((B, h, h, cin), device="cuda") * 0.5).to(torch.bfloat16)
    w = (torch.randn((cout, k * k * cin), device="cuda") * (k * k * cin) ** -0.5).to(torch.bfloat16)
    bias = torch.zeros(cout, device="cuda")
    out = torch.empty((B, h, h, cout), dtype=torch.bfloat16, device="cuda")
    stats = torch.zeros(2 * cout, device="cuda")
    res = torch.randn((B, h, h, cout), device="cuda").to(torch.bfloat16)
    flops = 2.0 * B * h * h * k * k * cin * cout
    byt = 2.0 * B * h * h * (cin + cout)
    for label, fn, extra in (
            ("fwd", lambda: ops.conv_gemm(x, w, bias=bias, ksize=k, relu=True, out=out), 0),
            ("fwd+stats", lambda: ops.conv_gemm(x, w, bias=bias, ksize=k, relu=True, stats=stats, out=out), 0),
            ("fwd+res", lambda: ops.conv_gemm(x, w, res1=res, ksize=k, relu=False, out=out), 2.0 * B * h * h * cout)):
        t = timeit(fn)
        print(f"{name:20s} {label:12s} {t * 1e6:9.1f} {flops / t / 1e12:9.1f} {(byt + extra) / t / 1e9:8.0f}")
    dy = torch.randn((B, h, h, cout), device="cuda").to(torch.bfloat16)
    dw = torch.zeros((cout, k * k * cin), device="cuda")
    t = timeit(lambda: ops.conv_wgrad(x, dy, ksize=k, dw=dw))
    print(f"{name:20s} {'wgrad':12s} {t * 1e6:9.1f} {flops / t / 1e12:9.1f} {byt / t / 1e9:8.0f}")
