"""Summarise an `ncu --set full` report (read with `ncu -i FILE --page raw --csv`) into one markdown table row per captured launch.
usage: python tools/ncu_summary.py gpurun_out/r02_targets.ncu-rep [gpurun_out/other.ncu-rep ...]"""
import csv, io, re, subprocess, sys

COLS = [("us", "gpu__time_duration.sum", 1e-3), ("DRAM rd GB", "dram__bytes_read.sum", None), ("DRAM wr GB", "dram__bytes_write.sum", None),
        ("DRAM %", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", 1), ("FMA-heavy %", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", 1),
        ("ALU %", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", 1), ("FP64 %", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", 1),
        ("issue %", "smsp__issue_active.avg.pct_of_peak_sustained_active", 1), ("warps %", "sm__warps_active.avg.pct_of_peak_sustained_active", 1),
        ("regs", "launch__registers_per_thread", 1), ("L2 hit %", "lts__t_sector_hit_rate.pct", 1),
        ("long_sb", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", 1), ("wait", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", 1),
        ("math_throttle", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", 1),
        ("tensor %", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 1)]
TO_GB = {"byte": 1e-9, "Kbyte": 1e-6, "Mbyte": 1e-3, "Gbyte": 1.0, "Tbyte": 1e3}
TO_US = {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "s": 1e6, "second": 1e6}


def main():
    print("| kernel | grid x block | " + " | ".join(c[0] for c in COLS) + " |\n|---|---|" + "---:|" * len(COLS))
    for path in sys.argv[1:]:
        raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        head, units = rows[0], rows[1]
        idx = {}
        for i, h in enumerate(head):
            idx.setdefault(h.split(".", 2)[-1] if h.count(".") > 3 else h, i); idx.setdefault(h, i)
        def col(name):
            for h, i in idx.items():
                if h.endswith(name):
                    return i
            return None
        for r in rows[2:]:
            name = re.sub(r"\(.*", "", r[head.index("Kernel Name")])
            cells = []
            for label, metric, scale in COLS:
                i = col(metric)
                if i is None or r[i] == "":
                    cells.append("-"); continue
                v = float(r[i].replace(",", ""))
                u = units[i]
                if label == "us": v *= TO_US.get(u, 1.0)
                elif label.startswith("DRAM") and label.endswith("GB"): v *= TO_GB.get(u, 1.0)
                cells.append(f"{v:.3f}" if abs(v) < 10 else f"{v:.1f}")
            print(f"| `{name}` | {r[head.index('Grid Size')]} x {r[head.index('Block Size')]} | " + " | ".join(cells) + " |")


if __name__ == "__main__":
    main()
