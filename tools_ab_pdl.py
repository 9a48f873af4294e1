import sys, torch
sys.path.insert(0, '.')
import hgb200
from hgb200 import ops, _lib
lib = _lib.lib
for B in (32, 128, 256):
    model = hgb200.HourglassModel(17, 8, 256, (256, 256, 3), "sigmoid", seed=1)
    model.compile(optimizer=hgb200.Adam(1e-3), loss=hgb200.loss.weighted_mse)
    img = torch.rand((B, 256, 256, 3), device="cuda")
    tg = ops.render_targets(torch.rand((B, 17), device="cuda") * 64, torch.rand((B, 17), device="cuda") * 64,
                            torch.randint(0, 3, (B, 17), device="cuda", dtype=torch.int32), 64, 64)
    for _ in range(3):
        model.train_step_device(img, tg)
    for rep in range(2):
        for flag in (0, 1):
            lib.hgb_debug_set(7, 2 if flag else 1)
            model.train_step_device(img, tg)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                model.train_step_device(img, tg)
            e1.record()
            torch.cuda.synchronize()
            print(f"B={B} pdl={'off' if flag else 'on '} {e0.elapsed_time(e1)/5:.2f} ms/step", flush=True)
    lib.hgb_debug_set(7, 0)
    del model, img, tg
    torch.cuda.empty_cache()
