"""A/B of the multi-stream lane schedule (hgb_debug_set(8, 1) = in-order replay on one stream) and of the
SM reserve of the skip lanes (key 9), plus the host cost of enqueueing one forward pass into an empty queue.
    python tools_lanes_ab.py [--batches 32,64,256]"""
import argparse, sys, time
import ctypes as C
import torch
sys.path.insert(0, '.')
import hgb200
from hgb200 import ops, _lib
ap = argparse.ArgumentParser()
ap.add_argument("--batches", default="32,64,128,256")
ap.add_argument("--reserves", default="20")
ap.add_argument("--wfirst", type=int, default=0)
ap.add_argument("--halo-min", type=int, default=0)
ap.add_argument("--single-store", type=int, default=0)
ap.add_argument("--pdl", type=int, default=0, help="0 adaptive, 1 force on, 2 force off")
ap.add_argument("--late", type=int, default=0, help="1 = early trigger (old behaviour)")
ap.add_argument("--bwd-pdl", type=int, default=0)
ap.add_argument("--wgrad-ctas", type=int, default=0)
ap.add_argument("--wgrad-prio", type=int, default=0)
ap.add_argument("--wide-min", type=int, default=0)
a = ap.parse_args()
lib = _lib.lib
lib.hgb_debug_set(10, a.wfirst)
lib.hgb_debug_set(15, a.halo_min)
lib.hgb_debug_set(16, a.single_store)
lib.hgb_debug_set(7, a.pdl)
lib.hgb_debug_set(18, a.late)
lib.hgb_debug_set(19, a.bwd_pdl)
lib.hgb_debug_set(20, a.wgrad_ctas)
lib.hgb_debug_set(21, a.wgrad_prio)
lib.hgb_debug_set(22, a.wide_min)
for B in [int(b) for b in a.batches.split(",")]:
    model = hgb200.HourglassModel(17, 8, 256, (256, 256, 3), "sigmoid", seed=1)
    model.compile(optimizer=hgb200.Adam(1e-3), loss=hgb200.loss.weighted_mse)
    img = torch.rand((B, 256, 256, 3), device="cuda")
    tg = ops.render_targets(torch.rand((B, 17), device="cuda") * 64, torch.rand((B, 17), device="cuda") * 64,
                            torch.randint(0, 3, (B, 17), device="cuda", dtype=torch.int32), 64, 64)
    for _ in range(3):
        model.train_step_device(img, tg)
    plan = model._plan(B, True)
    modes = [("in-order", 1, 0)] + [(f"lanes r{r}", 0, int(r)) for r in a.reserves.split(",")]
    for rep in range(2):
        for name, single, reserve in modes:
            lib.hgb_debug_set(8, single)
            lib.hgb_debug_set(9, reserve)
            model.train_step_device(img, tg)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                model.train_step_device(img, tg)
            e1.record()
            torch.cuda.synchronize()
            # host cost: forward only (fits the launch queue), queue empty at the start
            t0 = time.perf_counter()
            _lib.check(lib.hgb_model_forward(plan.handle, _lib.ptr(img), 1, None, _lib.stream_ptr()))
            t1 = time.perf_counter()
            torch.cuda.synchronize()
            t2 = time.perf_counter()
            print(f"B={B} {name:10s} {e0.elapsed_time(e1)/5:8.2f} ms/step | forward: host enqueue {1e3*(t1-t0):6.2f} ms, "
                  f"done after {1e3*(t2-t0):6.2f} ms", flush=True)
    lib.hgb_debug_set(8, 0); lib.hgb_debug_set(9, 0)
    del model, img, tg, plan
    torch.cuda.empty_cache()
