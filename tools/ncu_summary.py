"""Condenses `ncu --page raw --csv` exports into the per-kernel table kept under profiles/."""
import csv
import sys

COLS = [
    ("Kernel Name", "kernel"), ("gpu__time_duration.sum", "ms"), ("launch__grid_size", "grid"),
    ("launch__registers_per_thread", "regs"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps act %"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "fma cyc % (50 = IMAD peak)"),
    ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "alu cyc %"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "thr/inst"),
    ("dram__bytes_read.sum", "dram rd"), ("dram__bytes_write.sum", "dram wr"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "st math"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "st wait"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "st long_sb"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "st short_sb"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "st barrier"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "st lg_thr"),
    ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "st mio"),
]


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    cols = [(c, n) for c, n in COLS if c in idx]
    print("| " + " | ".join(n + (f" [{units[idx[c]]}]" if units[idx[c]] and c != "Kernel Name" else "") for c, n in cols) + " |")
    print("|" + "---|" * len(cols))
    for r in data:
        out = []
        for c, _ in cols:
            v = r[idx[c]]
            if c == "Kernel Name":
                v = v.split("(")[0].replace("void ", "")[:28]
            else:
                try:
                    v = f"{float(v.replace(',', '')):.4g}"
                except ValueError:
                    pass
            out.append(v)
        print("| " + " | ".join(out) + " |")


if __name__ == "__main__":
    main(sys.argv[1])
