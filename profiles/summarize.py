#!/usr/bin/env python
"""Summarise an .ncu-rep (brought back in gpurun_out/) into a small text file under profiles/.
usage: python profiles/summarize.py gpurun_out/prof.ncu-rep profiles/r1_pq_scan.txt"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed", "smsp__warps_eligible.avg.per_cycle_active",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers", "launch__grid_size", "launch__block_size",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.max"]


def main(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    lines = [f"# ncu --set full --clock-control none summary of {rep}", ""]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        lines.append(f"## {d.get('Kernel Name', '?')}  (id {d.get('ID', '?')})")
        for k in KEYS:
            if k in d:
                lines.append(f"{k:75s} {d[k]:>18s} {u[k]}")
        st = [(h, float(v.replace(',', ''))) for h, v in d.items()
              if h.startswith("smsp__pcsamp_warps_issue_stalled") and not h.endswith("_not_issued") and v]
        tot = sum(v for _, v in st) or 1.0
        lines.append("stall reasons (pc sampling): " + ", ".join(
            f"{h.replace('smsp__pcsamp_warps_issue_stalled_', '')} {v / tot * 100:.1f}%" for h, v in sorted(st, key=lambda x: -x[1])[:8]))
        lines.append("")
    open(out, "w").write("\n".join(lines))
    print("\n".join(lines))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
