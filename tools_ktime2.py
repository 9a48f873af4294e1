import ctypes as C, sys, torch
sys.path.insert(0, '.')
import hgb200
from hgb200 import ops, _lib
lib = _lib.lib
B=256
for cin, cout in ((128, 256), (256, 256), (256, 128)):
    x = (torch.randn((B, 64, 64, cin), device="cuda") * 0.5).to(torch.bfloat16)
    w = (torch.randn((cout, cin), device="cuda") * cin ** -0.5).to(torch.bfloat16)
    bias = torch.zeros(cout, device="cuda")
    out = torch.empty((B, 64, 64, cout), dtype=torch.bfloat16, device="cuda")
    stats = torch.zeros(2 * cout, device="cuda")
    for _ in range(5): ops.conv_gemm(x, w, bias=bias, ksize=1, relu=True, stats=stats, out=out)
    torch.cuda.synchronize()
    buf = (C.c_longlong * 32)()
    assert lib.hgb_debug_ktime(buf) == 0
    t = list(buf)
    f = lambda i: (t[i] - t[13]) / 1900.0
    print(f"{cin}->{cout}: epilogue tile 4: wait-start 0, acc-ready {f(14):.2f}, staged {f(15):.2f}, store-issued {f(16):.2f}, stats-done {f(17):.2f} | tile 5: wait-start {f(18):.2f}, acc-ready {f(19):.2f}, staged {f(20):.2f}, store-issued {f(21):.2f}, stats-done {f(22):.2f}")
    print(f"        MMA tile 3: tempty-wait {f(23):.2f}, got {f(24):.2f}, committed {f(25):.2f} | tile 4: tempty-wait {f(26):.2f}, got {f(27):.2f}, committed {f(28):.2f}")
