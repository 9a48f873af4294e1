"""Debug aid (NOT a pytest; lives under tests/ because it imports the oracle, which only test infrastructure may do): per-layer isolation of the CUDA forward pass.
For every conv: recompute it in fp32 torch FROM THE DEVICE'S OWN INPUT and compare with the device
output (isolated error), next to the error accumulated against the fp32 oracle."""
import ctypes as C
import sys

import numpy as np
import torch
import torch.nn.functional as F

sys.path.insert(0, ".")
import hgb200  # noqa: E402
from oracle import network_oracle as norc  # noqa: E402
from tests.test_gpu_network import _inputs  # noqa: E402

S, B = int(sys.argv[1]) if len(sys.argv) > 1 else 1, int(sys.argv[2]) if len(sys.argv) > 2 else 8
training = int(sys.argv[3]) if len(sys.argv) > 3 else 1
lib, chk = hgb200._lib.lib, hgb200._lib.check
images, targets = _inputs(B)
spec = norc.param_spec(17, S, 256)
weights = norc.init_params(spec, seed=2, perturb_bn=False)
model = hgb200.HourglassModel(17, S, 256, (256, 256, 3), "sigmoid")
model.set_weights_dict(weights)
plan = model._plan(B, True)
outs = model.forward_device(torch.as_tensor(images, device="cuda"), training=bool(training), plan=plan)
torch.cuda.synchronize()
_, _, taps = norc.forward(weights, images, 17, S, 256, training=bool(training), return_taps=True)


def fetch(i, which):
    chk(lib.hgb_debug_set(2, which))
    off, dims = C.c_int64(), (C.c_int * 4)()
    rc = lib.hgb_model_conv_output(plan.handle, i, C.byref(off), C.byref(dims))
    chk(lib.hgb_debug_set(2, 0))
    if rc:
        return None
    n = dims[0] * dims[1] * dims[2] * dims[3]
    return plan.arena[off.value:off.value + 2 * n].view(torch.bfloat16).float().reshape(*dims)


def errs(a, b):
    d = (a - b).abs()
    return d.max().item() / max(b.abs().max().item(), 1e-12), (d.pow(2).sum().sqrt() / b.pow(2).sum().sqrt().clamp_min(1e-12)).item()


print(f"{'conv':44s} {'iso max':>9s} {'iso l2':>9s} {'acc max':>9s} {'acc l2':>9s} {'bn iso max':>10s} {'bn iso l2':>9s}")
for i, c in enumerate(model.conv_table()):
    name = c["name"]
    y = fetch(i, 0)[..., :c["cout"]]
    ref = taps[name].detach().permute(0, 2, 3, 1).cuda()
    if "predict" in name and model.predict_activation == "sigmoid":
        y_cmp = torch.sigmoid(y)
    else:
        y_cmp = y
    acc = errs(y_cmp, ref)
    iso = (float("nan"), float("nan"))
    if c["k"] != 7:
        x = fetch(i, 1)[..., :c["cin"]]
        w = torch.as_tensor(weights[name + "/kernel"], device="cuda").to(torch.bfloat16).float().permute(3, 2, 0, 1)
        bias = torch.as_tensor(weights[name + "/bias"], device="cuda")
        r = F.conv2d(x.permute(0, 3, 1, 2), w, bias, padding=c["k"] // 2).permute(0, 2, 3, 1)
        lin = ("_conv_1x1_2" in name and name.startswith("hg")) or ("_conv_1x1_3" in name and name.startswith("hg") and "sample" not in name) or "predict" in name
        if not lin:
            r = torch.relu(r)
        if "predict" not in name and not (lin):
            iso = errs(y, r)
        elif "predict" in name:
            iso = errs(y, r)
    bniso = (float("nan"), float("nan"))
    z = fetch(i, 2)
    if z is not None and training and "conv_1x1_3" not in name:
        mean = y.mean(dim=(0, 1, 2))
        var = y.var(dim=(0, 1, 2), unbiased=False)
        zr = (y - mean) / torch.sqrt(var + 1e-3)
        bniso = errs(z, zr)
    print(f"{name:44s} {iso[0]:9.4f} {iso[1]:9.4f} {acc[0]:9.4f} {acc[1]:9.4f} {bniso[0]:10.4f} {bniso[1]:9.4f}")
for s in range(S):
    ref = None
print("heat", [errs(outs[s], torch.sigmoid(torch.zeros(1)).cuda() * 0 + taps[f"hg{s}_conv_1x1_predict"].detach().permute(0, 2, 3, 1).cuda()) for s in range(S)])
