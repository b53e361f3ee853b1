"""Summarise an `ncu --page raw --csv` export: one line per launch with the metrics the roofline discussion
needs (duration, tensor-pipe % of peak valid for tcgen05, DRAM / L2 bytes, hit rate, occupancy).
usage: python tools/ncu_summary.py gpurun_out/r2h_targets_raw.csv > profiles/r2_ncu_targets_summary.txt"""
import csv
import re
import sys

COLS = [
    ("us", "gpu__time_duration.sum", 1.0),
    ("tensor_ops%", "sm__ops_path_tensor_op_hmma_src_bf16_dst_fp32_sparsity_off.sum.pct_of_peak_sustained_elapsed", 1.0),
    ("tensor_act%", "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", 1.0),
    ("sm%", "sm__throughput.avg.pct_of_peak_sustained_elapsed", 1.0),
    ("dram_rd_MB", "dram__bytes_read.sum", 1.0),
    ("dram_wr_MB", "dram__bytes_write.sum", 1.0),
    ("dram%", "FBSP.TriageCompute.dram__throughput.avg.pct_of_peak_sustained_elapsed", 1.0),
    ("l2%", "lts__throughput.avg.pct_of_peak_sustained_elapsed", 1.0),
    ("l2_hit%", "lts__t_sector_hit_rate.pct", 1.0),
    ("warps%", "sm__warps_active.avg.pct_of_peak_sustained_active", 1.0),
    ("regs", "launch__registers_per_thread", 1.0),
]


def to_float(v, unit):
    try:
        x = float(v.replace(",", ""))
    except ValueError:
        return float("nan")
    scale = {"Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3, "byte": 1e-6, "ns": 1e-3, "us": 1.0, "ms": 1e3, "msecond": 1e3,
             "usecond": 1.0, "nsecond": 1e-3}
    return x * scale.get(unit, 1.0)


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    print(f"{'kernel':44s} {'grid':>6s} " + " ".join(f"{n:>11s}" for n, _, _ in COLS))
    for r in data:
        name = re.sub(r"^(void )?sisr::(<unnamed>|\(anonymous namespace\))::", "", r[idx["Kernel Name"]])
        name = re.sub(r"\(.*", "", name)[:44]
        grid = r[idx["Grid Size"]].replace(" ", "")
        grid = re.sub(r"[(),]", " ", grid).split()[0]
        vals = []
        for _, col, _ in COLS:
            i = idx.get(col)
            vals.append(to_float(r[i], units[i]) if i is not None else float("nan"))
        print(f"{name:44s} {grid:>6s} " + " ".join(f"{v:11.2f}" for v in vals))


if __name__ == "__main__":
    main(sys.argv[1])
