#!/usr/bin/env python
"""Summarise ONE kernel launch of an `ncu --set full` report as the JSON bench.py's `roofline.traffic` reads.

    ncu -i gpurun_out/prof_gemm.ncu-rep --page raw --csv > /tmp/raw.csv
    python tools/ncu_extract.py /tmp/raw.csv "umma_gemm_kernel<256, 0, 0, 0, 1>" 2 615776256 2*14144*3072*15360 \
        "capture description" > profiles/rNN_gemm_ncu_full.json

argv: raw csv, kernel-name substring, which matching launch (0-based), algorithmic bytes per launch, FLOPs per launch
(a Python expression), free-text description of the capture.
"""
import csv
import json
import sys

SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1e-3, "us": 1.0, "ms": 1e3,
         "s": 1e6}


def main():
    path, pat, which, alg_bytes, flops_expr, desc = sys.argv[1:7]
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    sel = [r for r in rows[2:] if pat in r[col["Kernel Name"]]]
    r = sel[int(which)]

    def val(name, to=1.0):
        i = col[name]
        return float(r[i].replace(",", "")) * SCALE.get(units[i], 1.0) / to

    dur_us = val("gpu__time_duration.sum")
    rd, wr = val("dram__bytes_read.sum"), val("dram__bytes_write.sum")
    flops = float(eval(flops_expr))
    out = {
        "capture": desc,
        "kernel": r[col["Kernel Name"]][:60],
        "launch_index_among_matches": int(which),
        "duration_us": dur_us,
        "dram_bytes_read": rd,
        "dram_bytes_write": wr,
        "traffic_bytes_per_launch": rd + wr,
        "algorithmic_bytes_per_launch": int(alg_bytes),
        "l2_to_sm_read_bytes": val("l1tex__m_xbar2l1tex_read_bytes.sum") if "l1tex__m_xbar2l1tex_read_bytes.sum" in col else None,
        "tensor_pipe_active_pct": None,
        "sm_cycles_active": float(r[col["TPC.TriageCompute.sm__cycles_active.avg"]]) if "TPC.TriageCompute.sm__cycles_active.avg" in col else None,
        "achieved_tflops": flops / (dur_us * 1e-6) / 1e12,
        "grid": r[col["Grid Size"]],
    }
    for name in ("sm__inst_executed_pipe_tensor_op_hmma.avg.pct_of_peak_sustained_active",
                 "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
                 "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active"):
        if name in col:
            out["tensor_pipe_active_pct"] = float(r[col[name]])
            out["tensor_pipe_metric"] = name
            break
    if out.get("sm_cycles_active"):
        out["sm_clock_ghz_during_capture"] = out["sm_cycles_active"] / (dur_us * 1e3)
    for name in ("launch__registers_per_thread", "launch__cluster_size", "launch__occupancy_limit_shared_mem"):
        if name in col:
            out[name] = r[col[name]]
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
