#!/usr/bin/env python
"""Run ONE op class of the real execution plan in isolation, bracketed by cudaProfilerStart/Stop, so that
    ncu --profile-from-start off --set full --clock-control none --import-source on -o gpurun_out/x \
        python tools_ncu_op.py --type B_DGRAD --k 1 --cin 256 --cout 128 --h 64
captures exactly that kernel on its real operands (1-stack plan at the given batch: a training forward, the loss and
every backward op before the selected one are replayed first).  Without ncu it prints the op's CUDA-event time
(L2 flushed between repetitions)."""
import argparse
import ctypes as C

import torch

import hgb200
from hgb200 import _lib, ops

NAMES = ["F_IM2COL", "F_CONV", "F_BN", "F_POOL", "F_UPADD", "F_HEAD", "B_BN_REDUCE", "B_BN_APPLY", "B_WGRAD", "B_DGRAD",
         "B_RELU_MASK", "B_COLSUM", "B_POOL", "B_UPADD", "B_HEAD"]
ap = argparse.ArgumentParser()
ap.add_argument("--type", default="B_DGRAD")
ap.add_argument("--k", type=int, default=1)
ap.add_argument("--cin", type=int, default=256)
ap.add_argument("--cout", type=int, default=128)
ap.add_argument("--c", type=int, default=0, help="channel count for non-conv ops")
ap.add_argument("--h", type=int, default=64)
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--stacks", type=int, default=1)
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--fused-stats", type=int, default=-1, help="B_DGRAD only: 1 = must carry a fused BN reduction, 0 = must not")
ap.add_argument("--debug", default="")
ap.add_argument("--input-bn", type=int, default=-1, help="F_CONV only: 1 = must read a deferred input BatchNorm (conv_1x1_3), 0 = must not")
ap.add_argument("--ab", default="", help="key=value: time every op a second time with this debug knob set (A/B in one process)")
ap.add_argument("--specs", default="", help="several ops in one process: TYPE:k:cin:cout:h[:fused] separated by commas "
                                            "(non-conv ops: TYPE:0:0:C:h); inputs are those left behind by one full step")
a = ap.parse_args()
lib, chk = _lib.lib, _lib.check
for kv in [x for x in a.debug.split(",") if x]:
    k, v = kv.split("=")
    lib.hgb_debug_set(int(k), int(v))
want = NAMES.index(a.type)
backward = 1 if want >= 6 else 0
S, B = a.stacks, a.batch
model = hgb200.HourglassModel(17, S, 256, (256, 256, 3), "sigmoid", seed=1)
model.compile(optimizer=hgb200.Adam(1e-3), loss=hgb200.loss.weighted_mse)
img = torch.rand((B, 256, 256, 3), device="cuda")
tg = ops.render_targets(torch.rand((B, 17), device="cuda") * 64, torch.rand((B, 17), device="cuda") * 64,
                        torch.randint(0, 3, (B, 17), device="cuda", dtype=torch.int32), 64, 64)
plan = model._plan(B, True)
h = plan.handle
sp = _lib.stream_ptr
lib.hgb_debug_set(8, 1)          # in-order replay on the caller's stream
model.forward_device(img, training=True, plan=plan)
losses = torch.zeros(S, dtype=torch.float64, device="cuda")
chk(lib.hgb_model_loss(h, 0, _lib.ptr(tg), 1.0 / (B * 64 * 64 * 17), _lib.ptr(losses), sp()))
torch.cuda.synchronize()
info, cinfo, coffs = (C.c_int * 8)(), (C.c_int * 8)(), (C.c_int64 * 2)()
off, dims = C.c_int64(), (C.c_int * 4)()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def matches(info):
    ty, conv, bn, a0 = info[0], info[1], info[2], info[3]
    if ty != want:
        return False
    if conv >= 0 and want in (1, 8, 9):
        lib.hgb_model_conv_detail(h, conv, C.byref(cinfo), C.byref(coffs))
        lib.hgb_model_act_info(h, a0, C.byref(off), C.byref(dims))
        ok = cinfo[0] == a.k and cinfo[2] == a.cin and cinfo[3] == a.cout and dims[1] == a.h
        if ok and want == 9 and a.fused_stats >= 0:
            ok = (bn >= 0) == bool(a.fused_stats)
        if ok and want == 1 and a.input_bn >= 0:
            ok = bool(info[7] & 0xffff) == bool(a.input_bn)
        return ok
    lib.hgb_model_act_info(h, a0, C.byref(off), C.byref(dims))
    return dims[1] == a.h and (a.c == 0 or dims[3] == a.c)


if a.specs:
    chk(lib.hgb_model_backward(h, 0, S + 1, sp()))          # every tensor of the plan now holds valid data
    torch.cuda.synchronize()
    abk, abv = (int(x) for x in a.ab.split("=")) if a.ab else (None, None)
    for spec in a.specs.split(","):
        f = spec.split(":")
        a.type, a.k, a.cin, a.cout, a.h = f[0], int(f[1]), int(f[2]), int(f[3]), int(f[4])
        a.c = a.cout if a.k == 0 else 0
        a.fused_stats = int(f[5]) if len(f) > 5 else -1
        want = NAMES.index(a.type)
        backward = 1 if want >= 6 else 0
        hit = None
        for seg in (range(S, -1, -1) if backward else range(S + 1)):
            for i in range(lib.hgb_model_num_ops(h, seg, backward)):
                chk(lib.hgb_model_op_info(h, seg, backward, i, C.byref(info)))
                if matches(info):
                    hit = (seg, i, tuple(info))
                    break
            if hit:
                break
        if not hit:
            print(f"{spec}: no matching op")
            continue
        res = []
        for setting in ([None] if abk is None else [None, abv]):
            if abk is not None:
                lib.hgb_debug_set(abk, 0 if setting is None else setting)
            ts = []
            for rep in range(a.reps + 1):
                flush.zero_()               # evicts the operands from L2 and keeps the queue busy while the launch is enqueued
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                chk(lib.hgb_model_run_op(h, hit[0], backward, hit[1], _lib.ptr(img), 1, sp()))
                e1.record()
                torch.cuda.synchronize()
                if rep:
                    ts.append(e0.elapsed_time(e1) * 1e3)
            res.append(min(ts))
        if abk is not None:
            lib.hgb_debug_set(abk, 0)
        print(f"{spec:40s} info {hit[2]}  us: " + " | ".join(f"{r:.1f}" for r in res) + (f"   (second: debug {a.ab})" if a.ab else ""))
    raise SystemExit(0)

found = False
segs = range(S, -1, -1) if backward else range(S + 1)
for seg in segs:
    if found:
        break
    for i in range(lib.hgb_model_num_ops(h, seg, backward)):
        chk(lib.hgb_model_op_info(h, seg, backward, i, C.byref(info)))
        if backward and not matches(info):
            chk(lib.hgb_model_run_op(h, seg, backward, i, _lib.ptr(img), 1, sp()))
            continue
        if not matches(info):
            continue
        torch.cuda.synchronize()
        times = []
        for rep in range(a.reps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            if rep == a.reps - 1:
                torch.cuda.synchronize()
                torch.cuda.profiler.start()
            e0.record()
            chk(lib.hgb_model_run_op(h, seg, backward, i, _lib.ptr(img), 1, sp()))
            e1.record()
            torch.cuda.synchronize()
            if rep == a.reps - 1:
                torch.cuda.profiler.stop()
            times.append(e0.elapsed_time(e1) * 1e3)
        print(f"{a.type} seg {seg} op {i} info {tuple(info)}: us per launch {[round(t, 1) for t in times]}")
        found = True
        break
if not found:
    raise SystemExit("no matching op in the plan")
