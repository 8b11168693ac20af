"""Key per-launch metrics of the first kernel in an ncu report: python profiles/ncu_key_metrics.py rep.ncu-rep"""
import csv, io, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__cycles_elapsed.avg",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__block_size", "launch__grid_size",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "smsp__average_warp_latency_per_inst_issued.ratio"]
for w in want:
    if w in hdr:
        i = hdr.index(w)
        print(f"{w:90s} {units[i]:14s} {vals[i]}")
for i, h in enumerate(hdr):
    if "warps_issue_stalled" in h and h.endswith("per_issue_active.ratio") and float(vals[i] or 0) > 0.05:
        print(f"{h:90s} {vals[i]}")
