#!/usr/bin/env python
"""BASELINE.json config 5 -- heatmap kernel sweep: target rendering, weighted-MSE loss+gradient and v2 decode at
64x64 and 128x128 x 17 joints, batch 64-4096, timed with CUDA events through the C ABI (direct ctypes calls on
preallocated buffers), reported as achieved HBM GB/s of the ALGORITHMIC bytes:
  render : B*H*W*K*4 written          loss : read y_true + read y_pred + write grad (all fp32)
  decode : B*H*W*K*sizeof(elem) read
Between launches a 256 MB buffer is rewritten and a second one read back, so inputs never sit in the 126 MB L2 and
the L2 is left CLEAN (after a pure write the cache holds ~100 MB of dirty lines whose write-back would be charged to
the kernel under test: +35 % on a 285 MB launch).
    python tools_heatmap_bench.py [--json]"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def sweep(batches=(64, 256, 1024, 4096), sizes=(64, 128), iters=5):
    import torch
    import hgb200  # noqa: F401
    from hgb200 import _lib
    from hgb200._lib import lib, check, ptr, stream_ptr
    K = 17
    flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device="cuda")
    flush2 = torch.zeros(64 * 1024 * 1024, dtype=torch.float32, device="cuda")
    sink = torch.zeros((), dtype=torch.float32, device="cuda")
    rows = []

    def timed(fn):
        fn(); torch.cuda.synchronize()
        tot = 0.0
        for _ in range(iters):
            flush.fill_(1.0)                      # evict the inputs from L2 ...
            sink.add_(flush2.sum())               # ... and the dirty lines of that write (leaves clean lines behind)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record()
            torch.cuda.synchronize()
            tot += e0.elapsed_time(e1)
        return tot / iters * 1e-3

    for H in sizes:
        for B in batches:
            n = B * H * H * K
            if n * 4 * 3 > 40e9:
                continue
            g = torch.Generator(device="cuda").manual_seed(B + H)
            kx = torch.rand((B, K), device="cuda", generator=g) * (H + 8) - 4
            ky = torch.rand((B, K), device="cuda", generator=g) * (H + 8) - 4
            kv = torch.randint(0, 3, (B, K), device="cuda", generator=g, dtype=torch.int32)
            yt = torch.empty((B, H, H, K), dtype=torch.float32, device="cuda")
            yp = torch.rand((B, H, H, K), device="cuda", generator=g)
            grad = torch.empty_like(yp)
            loss = torch.zeros((), dtype=torch.float64, device="cuda")
            ws = torch.empty(int(lib.hgb_loss_workspace_bytes(B, K)), dtype=torch.uint8, device="cuda")
            idx = torch.empty((B, K, 4), dtype=torch.int32, device="cuda")
            kp = torch.empty((B, K, 3), dtype=torch.float32, device="cuda")
            ypb = yp.to(torch.bfloat16)
            st = stream_ptr()
            t = timed(lambda: check(lib.hgb_render_targets(ptr(kx), ptr(ky), ptr(kv), B, H, H, K, ptr(yt), st)))
            rows.append(("render", H, B, t, n * 4))
            t = timed(lambda: check(lib.hgb_loss_fwd_bwd(0, ptr(yt), ptr(yp), _lib.F32, B, H, H, K, 1.0 / n, ptr(loss), ptr(grad),
                                                        _lib.F32, ptr(ws), st)))
            rows.append(("weighted_mse+grad", H, B, t, n * 12))
            t = timed(lambda: check(lib.hgb_decode(ptr(yp), _lib.F32, B, H, H, K, 1e-6, 2, ptr(idx), ptr(kp), st)))
            rows.append(("decode_v2 f32", H, B, t, n * 4))
            t = timed(lambda: check(lib.hgb_decode(ptr(ypb), _lib.BF16, B, H, H, K, 1e-6, 2, ptr(idx), ptr(kp), st)))
            rows.append(("decode_v2 bf16", H, B, t, n * 2))
            del yt, yp, grad, ypb
            torch.cuda.empty_cache()
    return rows


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--json", action="store_true")
    a = ap.parse_args()
    rows = sweep()
    if a.json:
        print(json.dumps([dict(kernel=k, hw=h, batch=b, us=t * 1e6, gbps=by / t / 1e9) for k, h, b, t, by in rows]))
    else:
        print(f"{'kernel':20s} {'HxW':>5s} {'batch':>6s} {'us':>10s} {'GB/s':>8s}")
        for k, h, b, t, by in rows:
            print(f"{k:20s} {h:5d} {b:6d} {t * 1e6:10.1f} {by / t / 1e9:8.0f}")
