"""How long does the host take to enqueue one training step vs how long the GPU takes to run it?"""
import sys, time, torch
sys.path.insert(0, '.')
import hgb200
from hgb200 import ops
for B in (32, 64):
    model = hgb200.HourglassModel(17, 8, 256, (256, 256, 3), "sigmoid", seed=1)
    model.compile(optimizer=hgb200.Adam(1e-3), loss=hgb200.loss.weighted_mse)
    img = torch.rand((B, 256, 256, 3), device="cuda")
    tg = ops.render_targets(torch.rand((B, 17), device="cuda") * 64, torch.rand((B, 17), device="cuda") * 64,
                            torch.randint(0, 3, (B, 17), device="cuda", dtype=torch.int32), 64, 64)
    for _ in range(3):
        model.train_step_device(img, tg)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        model.train_step_device(img, tg)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"B={B}: host enqueue {1e3*(t1-t0)/5:.2f} ms/step, total {1e3*(t2-t0)/5:.2f} ms/step")
