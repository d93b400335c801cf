"""Key metrics of every launch in an .ncu-rep (run where ncu is installed; no GPU needed).
usage: python profiles/extract_ncu.py gpurun_out/prof.ncu-rep"""
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum", "sm__cycles_elapsed.max",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}
for r in rows[2:]:
    print(f"== {r[col['Kernel Name']][:90]}  grid {r[col['Grid Size']]} block {r[col['Block Size']]}")
    for w in WANT:
        if w in col:
            print(f"   {w:75s} {r[col[w]]:>16s} {units[col[w]]}")
