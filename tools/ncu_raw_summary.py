"""Compact per-kernel table from `ncu -i X.ncu-rep --page raw --csv` (one row per profiled launch).

usage: python tools/ncu_raw_summary.py raw.csv [raw2.csv ...]
Columns: duration, DRAM read / written, pipe and issue utilisation, occupancy, registers, grid x block.
"""
import csv
import sys

COLS = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram_rd"),
    ("dram__bytes_write.sum", "dram_wr"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu%"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma%"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps%"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
]


def main(path):
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr, units = rows[hi], rows[hi + 1]
    kn = hdr.index("Kernel Name")
    have = [(m, s) for m, s in COLS if m in hdr]
    print(f"== {path}")
    print("  " + " | ".join(f"{s} [{units[hdr.index(m)]}]" if units[hdr.index(m)] else s for m, s in have))
    for r in rows[hi + 2:]:
        if len(r) <= kn:
            continue
        name = r[kn].replace("void ", "").replace("smb::", "")
        name = name[:name.index("(")] if "(" in name else name
        vals = []
        for m, _ in have:
            v = r[hdr.index(m)]
            try:
                vals.append(f"{float(v.replace(',', '')):.4g}")
            except ValueError:
                vals.append(v)
        print(f"  {name[:58]:58s} " + " | ".join(vals))


if __name__ == "__main__":
    for p in sys.argv[1:]:
        main(p)
