#!/usr/bin/env python
"""profiles/*.ncu-rep (ncu --set full, one launch each) -> profiles/ncu_kernel_metrics.json, the table bench.py copies into
`tensor_pipe_util_pct` and `roofline.traffic`: per op class the kernel, its duration and SM clock under ncu, DRAM bytes read /
written per launch, tensor-pipe and tensor-memory-pipe activity.  Runs where ncu is installed (no GPU needed):
    python tools_ncu_metrics.py"""
import csv
import io
import json
import os
import subprocess

ROOT = os.path.dirname(os.path.abspath(__file__))
REPORTS = {   # op class (hgb200/profiling.py naming) -> report
    "F_CONV k3 128->128 @64": "r02_conv3x3_fwd_b256_halo.ncu-rep",
    "B_DGRAD k3 128->128 @64 +bnstats": "r02_conv3x3_dgrad_b256_halo.ncu-rep",
    "B_WGRAD k3 128->128 @64": "r02_conv3x3_wgrad3_b256.ncu-rep",
    "B_WGRAD k1 256->256 @64": "r02_wgrad_k1_256to256.ncu-rep",
    "F_CONV k1 128->256 @64": "r02_fconv_k1_128to256_bn_stats.ncu-rep",
    "F_CONV k1 128->256 @64 (plain: no input BatchNorm, no statistics; round-2 start)": "r02_fconv_k1_128to256.ncu-rep",
    "B_DGRAD k1 256->128 @64 +bnapply +bnstats +res": "r02_dgrad_k1_256to128_bnb.ncu-rep",
    "B_DGRAD k1 128->256 @64 +bnapply +bnstats": "r02_dgrad_k1_128to256_bnb.ncu-rep",
    "decode_v2 f32 64x64 batch 1024": "r02_decode_f32_b1024_final.ncu-rep",
}
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "usecond": 1.0, "msecond": 1e3, "nsecond": 1e-3}


def read(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    def get(suffix):
        for h, u, v in zip(hdr, units, vals):
            if h.endswith(suffix):
                return float(v.replace(",", "")) * UNIT.get(u, 1.0)
        return None
    name = dict(zip(hdr, vals))["Kernel Name"].replace("void hgb::", "").split("(CUtensorMap")[0].split("(const")[0]
    rd, wr = get("dram__bytes_read.sum"), get("dram__bytes_write.sum")
    return {"kernel": name.replace("(int)", "").replace("(bool)", ""),
            "duration_us_under_ncu": round(get("gpu__time_duration.sum"), 1),
            "sm_clock_ghz_under_ncu": round(get("sm__cycles_elapsed.avg.per_second") or 0.0, 3),
            "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes_per_launch": rd + wr,
            "tensor_pipe_pct": round(get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed") or 0.0, 1),
            "tensor_mem_pipe_pct": round(get("sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed") or 0.0, 1),
            "source": "profiles/" + os.path.basename(path) + " (ncu --set full --clock-control none, one launch, batch 256)"}


table = {}
for op, rep in REPORTS.items():
    p = os.path.join(ROOT, "profiles", rep)
    if os.path.exists(p):
        table[op] = read(p)
        if not op.startswith(("F_CONV", "B_DGRAD", "B_WGRAD")):
            table[op]["tensor_pipe_pct"] = None
with open(os.path.join(ROOT, "profiles", "ncu_kernel_metrics.json"), "w") as f:
    json.dump(table, f, indent=1)
for k, v in table.items():
    print(f"{k:50s} {v['duration_us_under_ncu']:8.1f} us  DRAM {v['dram_bytes_per_launch'] / 1e6:8.1f} MB  tensor {v['tensor_pipe_pct']}")
