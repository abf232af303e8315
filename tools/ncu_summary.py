"""Key metrics of an `ncu --set full` report (one column per captured launch): python tools/ncu_summary.py X.ncu-rep"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "gpc__cycles_elapsed.avg.per_second", "launch__grid_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__cycles_active.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "lts__lts2xbar_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes.sum.per_second",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.sum"]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
names = [r[hdr.index("Kernel Name")][:60] for r in rows[2:]]
print("kernel:", " | ".join(names))
for k in KEYS:
    if k in hdr:
        i = hdr.index(k)
        print(f"{k:90s} [{units[i]:>12s}] " + " | ".join(r[i] for r in rows[2:]))
