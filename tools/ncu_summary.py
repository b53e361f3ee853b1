"""Condensed per-kernel summary of an ncu report (--set full): duration, DRAM bytes, DRAM / L2 / tensor
utilisation.  usage: python tools/ncu_summary.py report.ncu-rep"""
import csv, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg", "sm__cycles_elapsed.max",
        "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "sm__warps_active.avg.pct_of_peak_sustained_active"]
def find(name):
    for h in hdr:
        if h.endswith(name):
            return col[h]
    return None
for r in data:
    name = r[col["Kernel Name"]].split("(")[0].replace("void ", "").replace("sisr::<unnamed>::", "")
    print(name)
    for w in want:
        i = find(w)
        if i is not None:
            print(f"    {w:75s} {r[i]:>14s} {units[i]}")
