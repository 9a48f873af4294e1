"""Op classes of an execution plan and their ALGORITHMIC work: the FLOPs (2*MAC, convolutions only) and the bytes every
operand tensor is worth when it is moved exactly once (bf16 activations, fp32 heat maps / images).  Used by bench.py
(roofline of the class with the largest time share, per-class table) and tools_profile_step.py.  Pure bookkeeping over
the C ABI's plan introspection (hgb_model_op_info / act_info / conv_detail): no kernels, no oracle.
"""
from __future__ import annotations

import ctypes as C

from ._lib import lib

OP_NAMES = ["F_IM2COL", "F_CONV", "F_BN", "F_POOL", "F_UPADD", "F_HEAD", "B_BN_REDUCE", "B_BN_APPLY", "B_WGRAD", "B_DGRAD",
            "B_RELU_MASK", "B_COLSUM", "B_POOL", "B_UPADD", "B_HEAD", "F_DW", "B_DW_DGRAD", "B_DW_WGRAD"]
RIDGE_FLOP_PER_BYTE = 212.0     # measured bf16 peak / measured HBM bandwidth


class PlanInfo:
    def __init__(self, handle, num_classes=17):
        self.h, self.K = handle, num_classes
        self._off, self._dims = C.c_int64(), (C.c_int * 4)()
        self._cinfo, self._coffs = (C.c_int * 8)(), (C.c_int64 * 2)()

    def act(self, i):
        lib.hgb_model_act_info(self.h, i, C.byref(self._off), C.byref(self._dims))
        return tuple(self._dims)

    def conv(self, i):
        lib.hgb_model_conv_detail(self.h, i, C.byref(self._cinfo), C.byref(self._coffs))
        return dict(zip(("ksize", "taps", "cin", "cout", "cin_pad", "cout_pad", "relu", "has_dgrad"), self._cinfo))

    def classify(self, info, fused_bn=False):
        """info = the 8 ints of hgb_model_op_info -> (class key, flops, bytes, selector) where selector =
        (op_type, k, cin, cout, h) as hgb_model_profile_conv takes it."""
        ty, conv, bn, a0, a1, a2, a3, flag = info
        name = OP_NAMES[ty]
        el = lambda a: (lambda d: d[0] * d[1] * d[2] * d[3])(self.act(a)) if a >= 0 else 0     # noqa: E731
        if ty in (1, 8, 9) and conv >= 0:
            c = self.conv(conv)
            n, hh, ww, _ = self.act(a0)
            M = n * hh * ww
            flops = 2.0 * M * c["taps"] * c["cin"] * c["cout"]
            if ty == 1:
                byt = 2.0 * (M * (c["cin_pad"] + c["cout_pad"]) + el(a2) + el(a3))
            elif ty == 8:
                byt = 2.0 * M * (c["cin_pad"] + c["cout_pad"])
            else:   # dgrad: dp in, dx out, residuals, the y of a fused BatchNorm-backward REDUCTION; with the BatchNorm-backward
                #        APPLY fused in: dz and y in, dp out instead of dp in
                byt = 2.0 * (M * (c["cin_pad"] + c["cout_pad"] * (3 if fused_bn else 1)) + el(a2) + el(a3) + (el(flag - 1) if bn >= 0 else 0))
            key = f"{name} k{c['ksize']} {c['cin']}->{c['cout']} @{hh}"
            if ty == 9 and fused_bn:
                key += " +bnapply"
            if ty == 9 and bn >= 0:
                key += " +bnstats"
            if ty == 9 and (a2 >= 0 or a3 >= 0):
                key += " +res"
            return key, flops, byt, (ty, c["ksize"] if c["cin"] != 147 else 7, c["cin"] if c["cin"] != 147 else 3, c["cout"], hh)
        n, hh, ww, cc = self.act(a0)
        M = n * hh * ww
        hm = 4.0 * M * self.K
        if ty == 0:      # stem patches: fp32 image in, bf16 patches out
            byt = 4.0 * n * (2 * hh) * (2 * ww) * 3 + 2.0 * el(a0)
        elif ty == 2:    # y in, residual in, output out (+ the lower level of a fused upsample-add merge: a3)
            byt = 2.0 * (el(a0) + el(a1) + el(a2) + el(a3))
        elif ty == 3:
            byt = 2.0 * (el(a0) + el(a1))
        elif ty == 4:
            byt = 2.0 * (el(a0) + el(a1) + el(a2))
        elif ty == 5:
            byt = 2.0 * (el(a0) + el(a1)) + hm
        elif ty == 6:
            byt = 2.0 * (el(a0) + el(a1))
        elif ty == 7:
            byt = 2.0 * (el(a0) + el(a1) + el(a2))
        elif ty == 10:
            byt = 2.0 * 3 * el(a0)
        elif ty == 11:
            byt = 2.0 * el(a0)
        elif ty == 12:   # x, dy, dx (+ the accumulate read of dx) (+ the y of a folded BatchNorm-backward reduction: a3)
            byt = 2.0 * (el(a0) + el(a1) + el(a2) * (2 if flag else 1) + (el(a3) if bn >= 0 else 0))
        elif ty == 13:   # (+ the y of a folded BatchNorm-backward reduction: a2)
            byt = 2.0 * (el(a0) + el(a1) + (el(a2) if bn >= 0 else 0))
        elif ty in (15, 16, 17):   # mobile variant, depthwise stencil: tensor in, tensor out (+ residuals) / two tensors in
            byt = 2.0 * (el(a0) + el(a1) + (el(a2) + el(a3) if ty == 16 else 0))
        else:            # B_HEAD: loss gradient + heat map (fp32), re-injection gradient in, logits gradient out
            n, hh, ww, cc = self.act(a1)
            M = n * hh * ww
            byt = 2.0 * 4.0 * M * self.K + 2.0 * (el(a0) + el(a1))
        key = f"{name} C{cc} @{hh}"
        if ty == 2 and a3 >= 0:
            key += " +pool" if flag & 1 else " +upadd"
        if ty in (12, 13) and bn >= 0:
            key += " +bnstats"
        return key, 0.0, byt, (ty, 0, 0, cc, hh)


def bound_of(flops, byt):
    return "tensor" if byt > 0 and flops / byt > RIDGE_FLOP_PER_BYTE else "hbm"


def summarize(plan_handle, num_classes=17):
    """After a step run under hgb_model_profile_all: {class: dict(launches, ms, flops, bytes, selector)} and the total ms."""
    pi = PlanInfo(plan_handle, num_classes)
    info, ms, fb = (C.c_int * 8)(), C.c_double(), (C.c_int * 3)()
    agg, tot = {}, 0.0
    for i in range(lib.hgb_model_profile_count(plan_handle)):
        if lib.hgb_model_profile_op(plan_handle, i, C.byref(info), C.byref(ms)):
            continue
        lib.hgb_model_profile_op_fused(plan_handle, i, C.byref(fb))
        key, flops, byt, sel = pi.classify(tuple(info), fused_bn=fb[0] >= 0)
        r = agg.setdefault(key, dict(launches=0, ms=0.0, flops=0.0, bytes=0.0, selector=sel))
        r["launches"] += 1
        r["ms"] += ms.value
        r["flops"] += flops
        r["bytes"] += byt
        tot += ms.value
    return agg, tot


def class_table(agg, tot, peak_tflops, peak_gbps, top=None):
    """Rows sorted by time: share of the summed op time, achieved TFLOP/s and GB/s, the bounding roofline and its fraction."""
    rows = []
    for key, r in sorted(agg.items(), key=lambda kv: -kv[1]["ms"]):
        sec = r["ms"] * 1e-3
        tf = r["flops"] / sec / 1e12 if r["flops"] else 0.0
        gb = r["bytes"] / sec / 1e9 if r["bytes"] else 0.0
        bound = bound_of(r["flops"], r["bytes"])
        frac = tf / peak_tflops if bound == "tensor" else gb / peak_gbps
        rows.append(dict(op=key, launches=r["launches"], ms=round(r["ms"], 3), share=round(r["ms"] / tot, 4),
                         avg_us=round(1e3 * r["ms"] / r["launches"], 1), tflops=round(tf, 1), gbps=round(gb, 1), bound=bound,
                         frac=round(frac, 3), selector=r["selector"]))
    return rows[:top] if top else rows
