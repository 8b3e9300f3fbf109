"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time and launch count per kernel for
ONE training step (the window between two consecutive `adamw_kernel<bf16>` launches), as a share of the step.

    python tools/summarize_launches.py gpurun_out/launches.csv [step_index] > profiles/rNN_launches_summary.txt
"""
import csv
import re
import sys


def short(name: str) -> str:
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(.*$", "", name)
    name = name.replace("gh::", "")
    if name.startswith("at::") or "at::native" in name or "cutlass" in name or name.startswith("nchw") or "elementwise" in name:
        m = re.search(r"([A-Za-z_0-9]+)(<|$)", name.split("::")[-1])
        return "torch:" + (name[:70] if len(name) < 70 else name[:70] + "...")
    return name[:110]


def main():
    path = sys.argv[1]
    which = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    names = [r["Kernel Name"] for r in rows]
    t_ns = [float(r["Metric Value"].replace(",", "")) for r in rows]
    ends = [i for i, n in enumerate(names) if "adamw_kernel<__nv_bfloat16>" in n]
    if len(ends) <= which:
        which = len(ends) - 1
    lo, hi = ends[which - 1] + 1, ends[which] + 1
    # the fp32 group's adamw follows the bf16 one
    while hi < len(names) and "adamw_kernel" in names[hi]:
        hi += 1
    agg = {}
    for n, t in zip(names[lo:hi], t_ns[lo:hi]):
        k = short(n)
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += t
    total = sum(v[1] for v in agg.values())
    ours = sum(v[1] for k, v in agg.items() if not k.startswith("torch:"))
    print(f"# {path}: step window = launches [{lo}, {hi}) of {len(rows)}; {hi - lo} launches, "
          f"sum of kernel durations {total / 1e6:.2f} ms (cold-cache, serialised under ncu)")
    print(f"# our kernels: {ours / total * 100:.1f} % of the summed time; torch/ATen plumbing kernels: {100 - ours / total * 100:.1f} %")
    print(f"{'share':>7} {'ms':>9} {'launches':>8}  kernel")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{v[1] / total * 100:6.2f}% {v[1] / 1e6:9.3f} {v[0]:8d}  {k}")


if __name__ == "__main__":
    main()
