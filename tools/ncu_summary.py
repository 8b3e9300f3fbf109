"""Per-kernel table of an `ncu --set full` report (raw page as CSV):

    ncu -i rep.ncu-rep --page raw --csv > /tmp/raw.csv ; python tools/ncu_summary.py /tmp/raw.csv
"""
import csv
import sys

COLS = [("gpu__time_duration.sum", "us"), ("dram__bytes_read.sum", "rd MB"), ("dram__bytes_write.sum", "wr MB"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor %"),
        ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps %"),
        ("l1tex__m_xbar2l1tex_read_bytes.sum", "L2->SM GB"), ("launch__registers_per_thread", "regs"),
        ("launch__occupancy_limit_shared_mem", "occ smem"), ("launch__occupancy_limit_registers", "occ regs")]
SCALE = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3, "ns": 1e-3, "us": 1.0, "ms": 1e3}


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    print(f"{'id':>3s} {'kernel':44s} {'grid':16s} " + " ".join(f"{n:>9s}" for _, n in COLS))
    for r in rows[2:]:
        vals = []
        for k, n in COLS:
            if k not in col:
                vals.append("-")
                continue
            v = float(r[col[k]].replace(",", ""))
            u = units[col[k]]
            if "GB" in n:
                v = v * SCALE.get(u, 1.0) / 1e3
            elif u in SCALE:
                v = v * SCALE[u]
            vals.append(f"{v:9.2f}" if v < 1e5 else f"{v:9.0f}")
        name = r[col["Kernel Name"]].replace("void ", "").replace("gh::", "")[:44]
        print(f"{r[0]:>3s} {name:44s} {r[col['Grid Size']]:16s} " + " ".join(f"{v:>9s}" for v in vals))


if __name__ == "__main__":
    main()
