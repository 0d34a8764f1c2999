"""Selected metrics of one .ncu-rep (raw page) as text:  python tools/ncu_summary.py x.ncu-rep ["header line"]"""
import csv
import io
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__cluster_size", "launch__shared_mem_per_block_dynamic",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_op_dmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__average_warp_latency_per_inst_issued.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
]
rep = sys.argv[1]
if len(sys.argv) > 2:
    print(sys.argv[2])
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr, units = rows[0], rows[1]
for data in rows[2:]:
    d = dict(zip(hdr, data))
    u = dict(zip(hdr, units))
    print(f"kernel {d.get('Kernel Name', '?')[:100]}")
    for k in WANT:
        if k in d and d[k] != "":
            print(f"{k:80s} {u.get(k, ''):18s} {d[k]}")
    extra = [k for k in hdr if "pipe_tensor" in k and k not in WANT and d.get(k, "") not in ("", "0")]
    for k in extra[:8]:
        print(f"{k:80s} {u.get(k, ''):18s} {d[k]}")
