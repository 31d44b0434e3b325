"""Condenses an `.ncu-rep` (`ncu --set full`) into the few lines judged under profiles/: per captured launch the
duration, DRAM traffic, pipe utilisation (FP64 / DMMA tensor sub-pipe), occupancy and the top stall reasons.
Usage: python tools/ncu_summary.py report.ncu-rep > profiles/<name>.txt"""
import csv
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum",
    "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__cycles_elapsed.max",
    "smsp__cycles_active.avg",
]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    print(f"# source: {rep} (ncu --set full --clock-control none); one block per captured launch")
    for r in rows[2:]:
        print(f"kernel: {r[col['Kernel Name']][:140]}")
        for k in WANT:
            hits = [h for h in hdr if h.endswith(k)]
            for h in hits[:1]:
                print(f"  {k:84s} {r[col[h]]:>16s} {units[col[h]]}")
        stalls = []
        for h in hdr:
            if "smsp__average_warp" in h and "issue_stalled" in h and h.endswith("_per_warp_active.pct") is False and h.endswith(".ratio"):
                try:
                    stalls.append((float(r[col[h]]), h.split("issue_stalled_")[1].split("_per_")[0]))
                except ValueError:
                    pass
        stalls.sort(reverse=True)
        if stalls:
            print("  top stalls (warp latency cycles per issued instr): " + ", ".join(f"{n}={v:.2f}" for v, n in stalls[:5]))
        print()


if __name__ == "__main__":
    main()
