#!/usr/bin/env python
"""Benchmark of the stacked-hourglass hot path: 8-stack, 256-channel hourglass TRAINING step
(forward, weighted-MSE loss, backward, Keras-Adam) at 256x256 with 64x64x17 heat maps -- the
metric BASELINE.json quotes ("8-stack HG train img/s at 1/2/4/8 B200").

    python bench.py --gpus N --steps K --warmup W            # our arm (CUDA, one process per GPU)
    python bench.py --impl reference --steps K --warmup W    # the reference path on the host CPU

One JSON line on stdout (rank 0).  `value` = images/s with inputs resident in HBM, timed with CUDA
events, max over ranks; `e2e` = images/s through the public Python API with pinned host buffers
(H2D of images + keypoints, device-side target rendering, D2H of the losses) inside the timed region;
`roofline` = the dominant kernel (3x3 128->128 convolution at 64x64, 43.8 % of the FLOPs) timed live
with CUDA event pairs inside the step; `heatmap_kernels` = achieved HBM GB/s of target rendering, weighted-MSE
loss+gradient and v2 decode (BASELINE config 5 point: 64x64x17, batch 1024; full sweep: tools_heatmap_bench.py);
`cpu_baseline` = the fp32 CPU restatement of the reference (oracle/network_oracle.py; TensorFlow cannot be
installed here) on a bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

STACKS, CHANNELS, KPTS = 8, 256, 17
GLOBAL_BATCH = 256
FWD_GFLOP_PER_IMG = 68.870            # conv FLOPs (2*MAC), 8 stacks, forward


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--global-batch", type=int, default=GLOBAL_BATCH)
    ap.add_argument("--stacks", type=int, default=STACKS)
    ap.add_argument("--cpu-batch", type=int, default=2, help="images per step of the CPU arm (bounded sample)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def _oracle_color_augment():
    """CPU leg only: the numpy restatement of the reference's colour augmentation, timed beside the kernel."""
    from oracle import input_oracle
    return input_oracle.color_augment


# --------------------------------------------------------------------------------------- CPU arm
def cpu_train_steps(stacks, batch, steps, warmup):
    """The reference's training step (model/hourglass.py + loss.py weighted_mse + Keras Adam) as restated in
    oracle/network_oracle.py, fp32, all host threads.  Returns (images/s, threads)."""
    import numpy as np
    import torch
    from oracle import heatmap_oracle as horc
    from oracle import network_oracle as norc
    threads = len(os.sched_getaffinity(0))
    torch.set_num_threads(threads)
    rng = np.random.default_rng(0)
    images = rng.random((batch, 256, 256, 3), dtype=np.float32)
    kx = rng.uniform(-4, 68, (batch, KPTS)).astype(np.float32)
    ky = rng.uniform(-4, 68, (batch, KPTS)).astype(np.float32)
    kv = rng.integers(0, 3, (batch, KPTS))
    weights = norc.init_params(norc.param_spec(KPTS, stacks, CHANNELS), seed=2)
    slots = {k: (np.zeros_like(v), np.zeros_like(v)) for k, v in weights.items() if "moving_" not in k}
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        targets = horc.render_targets(kx, ky, kv, 64, 64)
        _outs, _losses, grads = norc.loss_and_grads(weights, images, targets, "weighted_mse", KPTS, stacks, CHANNELS)
        for k, g in grads.items():
            norc.adam_step(weights[k], g, slots[k][0], slots[k][1], it + 1)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    return batch * len(times) / sum(times), threads, sum(times) / len(times)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    ips, threads, sec = cpu_train_steps(args.stacks, args.cpu_batch, args.steps, max(args.warmup, 1))
    sample = f"{args.stacks}-stack fwd+weighted_MSE+bwd+Adam, batch {args.cpu_batch} per step, fp32 torch-CPU restatement of the reference"
    print(json.dumps({
        "impl": "reference", "metric": "hourglass_8stack_train_images_per_sec", "value": ips, "unit": "img/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": max(args.warmup, 1), "ms_per_step": sec * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.stacks}-stack 256ch hourglass training step, 256x256 -> 64x64x17, bounded CPU sample",
                   "global_batch": args.cpu_batch},
        "cpu_baseline": {"value": ips, "unit": "img/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": ips, "unit": "img/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


# --------------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index):
        self.rows, self.proc, self.dev = [], None, device_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.dev)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx = float(r[2])
                for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                    if r[col].lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------- our arm
def run_ours(args):
    import ctypes as C
    import numpy as np
    import torch
    import torch.distributed as dist
    import hgb200
    from hgb200 import _lib, ops, parallel

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    ar = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        ar = parallel.enable()
    if args.global_batch % world:
        raise SystemExit(f"global batch {args.global_batch} is not divisible by {world} ranks")
    B = args.global_batch // world                                   # strong scaling: the global batch is fixed
    lib = _lib.lib

    model = hgb200.HourglassModel(KPTS, args.stacks, CHANNELS, (256, 256, 3), "sigmoid", seed=1234)
    model.compile(optimizer=hgb200.Adam(1e-3), loss=hgb200.loss.weighted_mse)
    gen = torch.Generator(device="cuda").manual_seed(100 + rank)
    images = torch.rand((B, 256, 256, 3), device="cuda", generator=gen)
    kx = torch.rand((B, KPTS), device="cuda", generator=gen) * 72 - 4
    ky = torch.rand((B, KPTS), device="cuda", generator=gen) * 72 - 4
    kv = torch.randint(0, 3, (B, KPTS), device="cuda", generator=gen, dtype=torch.int32)
    targets = ops.render_targets(kx, ky, kv, 64, 64)
    plan = model._plan(B, True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        return model.train_step_device(images, targets, global_batch=args.global_batch, allreduce=ar)

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()

    # ---- timed region 1: inputs resident in HBM (activations of one step: tens of GB >> 126 MB L2)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    lib.hgb_model_profile_conv(plan.handle, 1, 1, 3, 128, 128, 64)      # forward 3x3 128->128 @ 64x64
    launches0 = lib.hgb_model_launch_count(plan.handle)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        losses = step()
    ev1.record()
    barrier()
    ms = torch.tensor([ev0.elapsed_time(ev1)], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    total_ms = ms.item()
    launches = int(lib.hgb_model_launch_count(plan.handle) - launches0)
    pm, pn, pf = C.c_double(), C.c_int(), C.c_double()
    lib.hgb_model_profile_read(plan.handle, C.byref(pm), C.byref(pn), C.byref(pf))
    lib.hgb_model_profile_conv(plan.handle, 0, 1, 3, 128, 128, 64)
    clocks = sampler.stop() if rank == 0 else None
    loss_val = float(losses.sum().item())

    # ---- timed region 2: end to end through the public API, pinned host inputs, losses read back every step
    h_img = images.cpu().pin_memory()
    h_kx, h_ky, h_kv = kx.cpu().pin_memory(), ky.cpu().pin_memory(), kv.cpu().pin_memory()
    model.train_on_keypoints(h_img, h_kx, h_ky, h_kv)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        out = model.train_on_keypoints(h_img, h_kx, h_ky, h_kv)
    barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    h2d = h_img.numel() * 4 + (h_kx.numel() + h_ky.numel() + h_kv.numel()) * 4
    d2h = 8 * args.stacks

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass
    peak_tf = float(peaks.get("bf16_tflops_sustained", 1400.0))
    peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained" if peaks else "fallback 1.4 PFLOP/s sustained (B200_PROFILING.md)"
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "top_kernel_traffic.json")) as f:
            traffic = json.load(f).get("dram_bytes_per_launch")
    except Exception:
        pass
    ach = (pf.value / max(pn.value, 1)) / (pm.value / max(pn.value, 1) * 1e-3) / 1e12 if pn.value else None
    value = args.global_batch * args.steps / (total_ms * 1e-3)
    line = {
        "metric": "hourglass_8stack_train_images_per_sec", "value": value, "unit": "img/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": f"{args.stacks}-stack 256ch hourglass training step (fwd + weighted_MSE + bwd + Adam), "
                               f"256x256x3 -> 64x64x17, random-init weights",
                   "global_batch": args.global_batch, "per_gpu_batch": B, "parallelism": f"dp{world}",
                   "l2_note": "per-step activation working set is tens of GB (>> 126 MB L2); no explicit flush needed",
                   "loss_last_step": loss_val,
                   "model_tflops_per_step": 3 * FWD_GFLOP_PER_IMG * args.global_batch / 1e3 * args.stacks / STACKS},
        "e2e": {"value": args.global_batch * args.steps / e2e_s.item(), "unit": "img/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "api": "HourglassModel.train_on_keypoints(pinned images, kps_x, kps_y, kps_v)"},
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": {"bound": "tensor", "kernel": "conv_gemm_kernel<128,1,4,1,HALO> forward 3x3 128->128 @64x64",
                     "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s", "frac": (ach / peak_tf) if ach else None,
                     "traffic": traffic, "launches_timed": pn.value, "avg_launch_ms": pm.value / max(pn.value, 1),
                     "peak_source": peak_src},
        "model_flops_utilization": 3 * FWD_GFLOP_PER_IMG * args.stacks / STACKS * value / 1e3 / world / peak_tf,
    }
    if world == 1:
        try:     # BASELINE.json metric: "decode GB/s" (+ the other heat-map kernels), algorithmic bytes / CUDA-event time
            import tools_heatmap_bench
            peak_bw = float(peaks.get("hbm_gbs", 6500.0))
            line["heatmap_kernels"] = {
                "shape": "64x64x17, batch 1024, inputs evicted from L2 between launches",
                "peak_gbps": peak_bw,
                "kernels": {k: {"us": round(t * 1e6, 1), "gbps": round(by / t / 1e9, 1), "frac": round(by / t / 1e9 / peak_bw, 3)}
                            for k, _h, _b, t, by in tools_heatmap_bench.sweep(batches=(1024,), sizes=(64,), iters=5)}}
        except Exception as ex:   # the sweep is auxiliary: never lose the headline line over it
            line["heatmap_kernels"] = {"error": f"{type(ex).__name__}: {ex}"}
    if world == 1:
        try:     # the input path (decode / resize / augment) at batch 256, next to the reference's CPU ops
            import tools_input_bench
            peak_bw = float(peaks.get("hbm_gbs", 6500.0))
            line["input_kernels"] = {
                "shape": "batch 256 of 256x256x3 f32 (crop_resize from 640x480 u8), L2 flushed between launches; cpu = the "
                         "reference's per-example op (cv2 / numpy restatement) on one host thread",
                "peak_gbps": peak_bw,
                "kernels": {k: {kk: (round(v, 6) if isinstance(v, float) else v) for kk, v in
                                {"us": r["s"] * 1e6, "gbps": r.get("gbps"), "frac": (r["gbps"] / peak_bw) if "gbps" in r and not r.get("host_timed") else None,
                                 "images_per_s": r.get("gpu_images_per_s", r.get("images_per_s")),
                                 "cpu_images_per_s": r.get("cpu_images_per_s")}.items() if v is not None}
                            for k, r in tools_input_bench.sweep(batch=256, iters=3, cpu_color_fn=_oracle_color_augment()).items()}}
        except Exception as ex:
            line["input_kernels"] = {"error": f"{type(ex).__name__}: {ex}"}
    if world == 1 and not args.no_cpu_baseline:
        ips, threads, sec = cpu_train_steps(args.stacks, args.cpu_batch, 2, 1)
        line["cpu_baseline"] = {"value": ips, "unit": "img/s", "cores": threads, "kind": "port",
                                "sample": f"{args.stacks}-stack training step, batch {args.cpu_batch}, 1 warm-up + 2 timed steps, "
                                          f"fp32 torch-CPU restatement of the reference ({sec:.1f} s/step)"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
