#!/usr/bin/env python
"""ncu `--page raw --csv` export -> markdown table for profiles/.   usage: ncu_summary.py raw.csv [out.md] [title]"""
import csv
import json
import sys


def num(x):
    try:
        return float(x.replace(",", ""))
    except Exception:
        return None


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}

    def col(r, name):
        return r[idx[name]] if name in idx else ""

    def to_bytes(r, name):
        v, u = num(col(r, name)), units[idx[name]] if name in idx else ""
        if v is None:
            return None
        return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(u, 1.0)

    def to_ms(r, name):
        v, u = num(col(r, name)), units[idx[name]] if name in idx else ""
        if v is None:
            return None
        return v * {"s": 1e3, "ms": 1.0, "us": 1e-3, "ns": 1e-6}.get(u, 1.0)

    out = []
    title = sys.argv[3] if len(sys.argv) > 3 else sys.argv[1]
    out.append("# %s\n" % title)
    out.append("| kernel | grid x block | regs | time ms | DRAM rd GB | DRAM wr GB | DRAM GB/s | DRAM %% of ncu peak | SM %% | inst/elem-byte proxy (M warp-inst) | top stalls |")
    out.append("|---|---|---|---|---|---|---|---|---|---|---|")
    js = []
    for r in data:
        name = col(r, "Kernel Name")
        short = name.replace("void ", "").split("(")[0]
        t = to_ms(r, "gpu__time_duration.sum")
        rd, wr = to_bytes(r, "dram__bytes_read.sum"), to_bytes(r, "dram__bytes_write.sum")
        st = {}
        for h in hdr:
            if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
                v = num(r[idx[h]])
                if v is not None:
                    st[h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]] = v
        tot = sum(st.values()) or 1.0
        top = ", ".join("%s %.0f%%" % (k, 100 * v / tot) for k, v in sorted(st.items(), key=lambda x: -x[1])[:3])
        gbs = (rd + wr) / (t * 1e-3) / 1e9 if (t and rd is not None and wr is not None) else None
        out.append("| `%s` | %s x %s | %s | %.4f | %.3f | %.3f | %s | %s | %s | %s | %s |" % (
            short[:90], col(r, "Grid Size"), col(r, "Block Size"), col(r, "launch__registers_per_thread"), t or 0,
            (rd or 0) / 1e9, (wr or 0) / 1e9, "%.0f" % gbs if gbs else "", col(r, "dram__throughput.avg.pct_of_peak_sustained_elapsed")[:5],
            col(r, "sm__throughput.avg.pct_of_peak_sustained_elapsed")[:5],
            "%.1f" % ((num(col(r, "smsp__inst_executed.sum")) or 0) / 1e6), top))
        js.append({"kernel": short, "time_ms": t, "dram_read_bytes": rd, "dram_write_bytes": wr})
    text = "\n".join(out) + "\n"
    if len(sys.argv) > 2:
        open(sys.argv[2], "w").write(text)
        json.dump(js, open(sys.argv[2].replace(".md", ".json"), "w"), indent=1)
    else:
        print(text)


if __name__ == "__main__":
    main()
