#!/usr/bin/env python
"""Transpose `ncu -i REP --page raw --csv` into the `metric,unit,value` summaries kept under profiles/ (CPU only).

    python scripts/ncu_summary.py gpurun_out/r2_step.ncu-rep [launch index | kernel-name substring] > profiles/r2_step_kernel_ncu_full.csv
"""
import csv
import io
import subprocess
import sys

KEEP = (
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "launch__block_size", "launch__grid_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "launch__shared_mem_per_block_static", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "launch__occupancy_limit_warps", "launch__waves_per_multiprocessor",
    "sm__cycles_elapsed.max", "sm__cycles_active.avg", "sm__cycles_active.max", "sm__cycles_active.min",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.per_cycle_active",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.avg.per_cycle_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_subpipe_umma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_active", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
)


def main():
    rep = sys.argv[1]
    sel = sys.argv[2] if len(sys.argv) > 2 else "0"
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    names = [r[hdr.index("Kernel Name")] for r in data]
    if sel.isdigit():
        k = int(sel)
    else:
        k = next(i for i, n in enumerate(names) if sel in n)
    row = data[k]
    print(f"# {names[k]} -- launch {k} of {rep.split('/')[-1]} ({len(data)} profiled launches: {', '.join(sorted(set(n.split('(')[0] for n in names)))})")
    print("# ncu --set full --clock-control none --import-source on; per-launch values (cold-cache, serialised: no PDL overlap)")
    stall = []
    for h, u, v in zip(hdr, units, row):
        if h in KEEP:
            print(f"{h},{u},{v}")
        elif h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") or \
                (h.startswith("smsp__average_warp_latency_issue_stalled_") and h.endswith(".ratio")):
            try:
                stall.append((float(v.replace(",", "")), h))
            except ValueError:
                pass
    for v, h in sorted(stall, reverse=True)[:10]:
        print(f"{h},ratio,{v}")


if __name__ == "__main__":
    main()
