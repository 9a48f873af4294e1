"""In-kernel timeline of the persistent conv GEMM at small sizes (needs libhgb200 built with -DHGB_KTIME)."""
import ctypes as C, sys, torch
sys.path.insert(0, '.')
import hgb200
from hgb200 import ops, _lib
lib = _lib.lib
NAMES = ["entry", "setup done", "pdl_wait done", "syncthreads", "first stage landed", "last MMA committed", "epi: acc ready",
         "epi: tile staged", "epi: store issued", "epi: loop done", "epi: store read", "epi: stats flushed", "exit"]
for (name, k, cin, cout, h, B, stats) in [("k1 128->256 @4", 1, 128, 256, 4, 32, True), ("k3 128->128 @4", 3, 128, 128, 4, 32, True),
                                          ("k3 128->128 @16", 3, 128, 128, 16, 32, True), ("k1 256->128 @16", 1, 256, 128, 16, 32, False)]:
    x = (torch.randn((B, h, h, cin), device="cuda") * 0.5).to(torch.bfloat16)
    w = (torch.randn((cout, k * k * cin), device="cuda") * (k * k * cin) ** -0.5).to(torch.bfloat16)
    bias = torch.zeros(cout, device="cuda")
    out = torch.empty((B, h, h, cout), dtype=torch.bfloat16, device="cuda")
    st = torch.zeros(2 * cout, device="cuda") if stats else None
    for _ in range(5):
        ops.conv_gemm(x, w, bias=bias, ksize=k, relu=True, stats=st, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        ops.conv_gemm(x, w, bias=bias, ksize=k, relu=True, stats=st, out=out)
    e1.record()
    torch.cuda.synchronize()
    buf = (C.c_longlong * 32)()
    assert lib.hgb_debug_ktime(buf) == 0
    t = list(buf)[:13]
    print(f"{name} B={B}: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us per back-to-back launch; CTA 0 timeline (us at 1.9 GHz, from entry):")
    for n, v in zip(NAMES, t):
        print(f"    {n:22s} {(v - t[0]) / 1900.0:7.2f}")
