"""Key metrics of an .ncu-rep (read with `ncu -i ... --page raw --csv`) as a small text table.
usage: python scripts/ncu_summary.py gpurun_out/x.ncu-rep > profiles/rNN_x_summary.txt"""
import csv
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__shared_mem_per_block_dynamic", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__cycles_active.avg", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "lts__t_bytes.sum", "l1tex__t_bytes.sum", "lts__t_sector_hit_rate.pct",
]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
ki = hdr.index("Kernel Name")
print("# " + sys.argv[1] + "  (ncu --set full --clock-control none)")
for n, row in enumerate(rows[2:]):
    print(f"launch {n}: {row[ki][:120]}")
    for i, h in enumerate(hdr):
        if any(h == w or (h.startswith(w) and h[len(w):len(w) + 1] in ("", ".")) for w in WANT):
            if h.endswith(".per_second") or h.endswith("pct_of_peak_sustained_elapsed") and h.startswith("dram__bytes"):
                continue
            print(f"    {h:70s} {row[i]:>16s} {units[i]}")
