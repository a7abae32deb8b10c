"""Extracts a text summary from an .ncu-rep (raw metrics of the first kernel + the hottest SASS
lines with their stall-sample counts) for profiles/."""
import csv, subprocess, sys, io
rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, vals = rows[0], rows[2] if len(rows) > 2 else rows[1]
keys = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_static", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__average_warp_latency_per_inst_issued.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
units = rows[1]
with open(out, "w") as f:
    f.write("# %s\n" % rep)
    for h, u, v in zip(hdr, units, vals):
        if h in keys:
            f.write("%s = %s %s\n" % (h, v, u))
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(io.StringIO(src)))
    h2, data = r[1], [x for x in r[2:] if len(x) == len(r[1])]
    isrc, isamp, iex = h2.index("Source"), h2.index("# Samples"), h2.index("Instructions Executed")
    tot = sum(int(x[isamp]) for x in data)
    f.write("\n# hottest SASS instructions (stall samples, share of %d, times executed)\n" % tot)
    for x in sorted(data, key=lambda x: -int(x[isamp]))[:40]:
        f.write("%7s %5.1f%% %10s  %s\n" % (x[isamp], 100.0 * int(x[isamp]) / tot, x[iex], x[isrc].strip()[:100]))
print("wrote", out)
