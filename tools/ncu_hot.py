"""Hot SASS instructions of one kernel of an `ncu --set full --import-source on` report.

    ncu -i rep.ncu-rep --page source --csv --kernel-id :::N > /tmp/src.csv ; python tools/ncu_hot.py /tmp/src.csv [top]
"""
import csv
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    print("#", rows[0][1][:110])
    hdr = rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    data = [r for r in rows[2:] if len(r) > idx["# Samples"] and r[idx["# Samples"]].isdigit()]
    tot = sum(int(r[idx["# Samples"]]) for r in data)
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    agg = {h: sum(int(r[idx[h]]) for r in data if r[idx[h]].isdigit()) for h in stalls}
    print(f"# {len(data)} SASS instructions, {tot} warp-stall samples; by reason:",
          ", ".join(f"{k[6:]} {100 * v / tot:.1f}%" for v, k in sorted(((v, k) for k, v in agg.items()), reverse=True)[:8]))
    print("# samples  share  executed  instruction                                                   top stall")
    for r in sorted(data, key=lambda r: -int(r[idx["# Samples"]]))[:top]:
        s = int(r[idx["# Samples"]])
        st = sorted(((int(r[idx[h]]) if r[idx[h]].isdigit() else 0, h) for h in stalls), reverse=True)[:1]
        print(f"{s:8d} {100 * s / tot:5.1f}% {r[idx['Instructions Executed']]:>9s}  {r[idx['Source']].strip()[:62]:62s} {st[0][1][6:]}={st[0][0]}")


if __name__ == "__main__":
    main()
