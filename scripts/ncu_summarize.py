#!/usr/bin/env python3
"""
Summarise `ncu --page raw --csv` exports (scripts/gpu_round_check.sh) into a markdown table and a
column-filtered CSV for profiles/.

    python scripts/ncu_summarize.py gpurun_out/r01b_ncu_full_c4.csv profiles/r01b_ncu_full_c4.csv
"""
import csv
import sys

KEEP = [
    "ID", "Kernel Name", "Block Size", "Grid Size", "gpu__time_duration.sum", "dram__bytes_read.sum",
    "dram__bytes_write.sum", "dram__bytes_read.sum.per_second", "dram__bytes_write.sum.per_second",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__cycles_active.avg", "sm__cycles_elapsed.max", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sectors_op_red.sum", "lts__t_sectors_op_atom.sum",
]


def fnum(x):
    try:
        return float(x.replace(",", ""))
    except Exception:
        return float("nan")


def main():
    src, dst = sys.argv[1], sys.argv[2]
    rows = list(csv.reader(open(src)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    cols = [c for c in KEEP if c in idx]
    with open(dst, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(cols)
        w.writerow([units[idx[c]] for c in cols])
        for r in data:
            w.writerow([r[idx[c]] for c in cols])
    # markdown summary grouped by kernel name
    groups = {}
    for r in data:
        name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "")
        groups.setdefault((name, r[idx["Grid Size"]]), []).append(r)

    def unit_scale(col, target):
        u = units[idx[col]]
        table = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}
        return table.get(u, 1.0)

    print("| kernel | grid | n | time us | DRAM rd MB | DRAM wr MB | regs | SM % | L1 % | L2 % | DRAM % | warps active % | L2 hit % | warp inst |")
    print("|---|---|---|---|---|---|---|---|---|---|---|---|---|---|")
    for (name, grid), rs in groups.items():
        def avg(col):
            vals = [fnum(r[idx[col]]) for r in rs] if col in idx else [float("nan")]
            return sum(vals) / len(vals)
        t = avg("gpu__time_duration.sum") * unit_scale("gpu__time_duration.sum", "us")
        rd = avg("dram__bytes_read.sum") * unit_scale("dram__bytes_read.sum", "MB")
        wr = avg("dram__bytes_write.sum") * unit_scale("dram__bytes_write.sum", "MB")
        print(f"| `{name}` | {grid} | {len(rs)} | {t:.2f} | {rd:.2f} | {wr:.2f} | {avg('launch__registers_per_thread'):.0f} | "
              f"{avg('sm__throughput.avg.pct_of_peak_sustained_elapsed'):.1f} | {avg('l1tex__throughput.avg.pct_of_peak_sustained_elapsed'):.1f} | "
              f"{avg('lts__throughput.avg.pct_of_peak_sustained_elapsed'):.1f} | {avg('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):.1f} | "
              f"{avg('sm__warps_active.avg.pct_of_peak_sustained_active'):.1f} | {avg('lts__t_sector_hit_rate.pct'):.1f} | {avg('smsp__inst_executed.sum'):.3g} |")


if __name__ == "__main__":
    main()
