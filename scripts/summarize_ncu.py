#!/usr/bin/env python3
"""Summarise an .ncu-rep (read here, no GPU needed) into a small CSV/markdown for profiles/.

    python scripts/summarize_ncu.py gpurun_out/k1_full.ncu-rep profiles/r01_k1_full
"""
import csv
import subprocess
import sys

KEYS = [
    "Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_xu_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fmalite.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__cycles_elapsed.avg.per_second",
    "l1tex__m_l1tex2xbar_write_bytes.sum.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    keys = [k for k in KEYS if k in idx]
    with open(out + ".csv", "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(keys)
        w.writerow([units[idx[k]] for k in keys])
        for r in data:
            w.writerow([r[idx[k]] for k in keys])
    with open(out + ".md", "w") as f:
        f.write(f"# ncu summary of `{rep}` ({len(data)} launches, --set full --clock-control none)\n\n")
        f.write("| " + " | ".join(["#", "kernel", "grid", "time", "DRAM rd", "DRAM wr", "DRAM GB/s", "DRAM %", "SM %", "regs", "warps act %", "issue %", "tensor pipe %", "SM GHz"]) + " |\n")
        f.write("|" + "---|" * 14 + "\n")
        for i, r in enumerate(data):
            g = lambda k: r[idx[k]] if k in idx else ""
            u = lambda k: units[idx[k]] if k in idx else ""
            t = float(g("gpu__time_duration.sum").replace(",", ""))
            tu = u("gpu__time_duration.sum")
            tsec = t * {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}.get(tu, 1e-9)
            def byts(k):
                v = float(g(k).replace(",", ""))
                return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u(k), 1)
            rd, wr = byts("dram__bytes_read.sum"), byts("dram__bytes_write.sum")
            name = g("Kernel Name").split("(")[0][:40]
            f.write(f"| {i} | {name} | {g('Grid Size')} | {t:.1f} {tu} | {rd/1e6:.1f} MB | {wr/1e6:.1f} MB | {(rd+wr)/tsec/1e9:.0f} | "
                    f"{g('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed')} | {g('sm__throughput.avg.pct_of_peak_sustained_elapsed')} | "
                    f"{g('launch__registers_per_thread')} | {g('sm__warps_active.avg.pct_of_peak_sustained_active')} | "
                    f"{g('smsp__issue_active.avg.pct_of_peak_sustained_active')} | "
                    f"{g('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active')} | {g('sm__cycles_elapsed.avg.per_second')} |\n")
    print("wrote", out + ".csv", out + ".md")


if __name__ == "__main__":
    main()
