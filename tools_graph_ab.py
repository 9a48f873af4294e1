"""Does replaying the whole training step as ONE CUDA graph (captured through torch, lanes included) beat
stream launches?  python tools_graph_ab.py [--batches 32,64]"""
import argparse, sys
import torch
sys.path.insert(0, '.')
import hgb200
from hgb200 import ops, _lib
ap = argparse.ArgumentParser()
ap.add_argument("--batches", default="32,64,256")
a = ap.parse_args()
lib = _lib.lib
for B in [int(b) for b in a.batches.split(",")]:
    model = hgb200.HourglassModel(17, 8, 256, (256, 256, 3), "sigmoid", seed=1)
    model.compile(optimizer=hgb200.Adam(1e-3), loss=hgb200.loss.weighted_mse)
    img = torch.rand((B, 256, 256, 3), device="cuda")
    tg = ops.render_targets(torch.rand((B, 17), device="cuda") * 64, torch.rand((B, 17), device="cuda") * 64,
                            torch.randint(0, 3, (B, 17), device="cuda", dtype=torch.int32), 64, 64)
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(3):
            model.train_step_device(img, tg)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            model.train_step_device(img, tg)
        e1.record()
        torch.cuda.synchronize()
        print(f"B={B} stream launches: {e0.elapsed_time(e1) / 5:.2f} ms/step", flush=True)
        try:
            g = torch.cuda.CUDAGraph()
            it0 = model.optimizer.iterations
            with torch.cuda.graph(g, stream=s, capture_error_mode="relaxed"):
                losses = model.train_step_device(img, tg)
            for _ in range(2):
                g.replay()
            torch.cuda.synchronize()
            e0.record()
            for _ in range(5):
                g.replay()
            e1.record()
            torch.cuda.synchronize()
            print(f"B={B} one CUDA graph : {e0.elapsed_time(e1) / 5:.2f} ms/step  (losses {losses.sum().item():.4f})", flush=True)
        except Exception as ex:
            print(f"B={B} graph capture failed: {type(ex).__name__}: {ex}", flush=True)
    del model, img, tg
    torch.cuda.empty_cache()
