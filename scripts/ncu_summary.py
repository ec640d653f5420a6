"""Summarise an .ncu-rep (read here, no GPU needed) into a small text file for profiles/.
usage: python scripts/ncu_summary.py gpurun_out/x.ncu-rep profiles/x.txt"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "sm__cycles_elapsed.avg",
        "lts__throughput.avg.pct", "l1tex__data_pipe_lsu_wavefronts.avg.pct", "sm__inst_executed.sum",
        "smsp__issue_active.avg.pct", "sm__throughput.avg.pct", "launch__shared_mem_per_block_dynamic",
        "smsp__warp_issue_stalled", "dram__cycles_active.avg.pct"]
rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
lines = [f"# ncu --set full --clock-control none summary of {rep}"]
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
    lines.append(f"## kernel {name}")
    for h, u, v in zip(hdr, units, r):
        if any(k in h for k in KEYS) and ".max" not in h and ".min" not in h and ".sum.p" not in h:
            lines.append(f"{h} [{u}] = {v}")
open(out, "w").write("\n".join(lines) + "\n")
print("\n".join(lines[:60]))
