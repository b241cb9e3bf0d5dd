"""Compact per-kernel summary of an .ncu-rep (run here, no GPU needed):
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/ncu_<round>_<what>.txt"""
import csv
import io
import subprocess
import sys

METRICS = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram_rd"),
    ("dram__bytes_write.sum", "dram_wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor%"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps%"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__shared_mem_per_block_dynamic", "dsmem"),
]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    print(f"# {path}: {len(data)} kernel launches (ncu --set full --clock-control none; cold-cache, serialised)")
    for r in data:
        name = r[idx["Kernel Name"]].replace("void ", "").replace("<unnamed>::", "")[:58]
        parts = [f"{name:58s}"]
        for m, short in METRICS:
            if m in idx:
                parts.append(f"{short}={r[idx[m]]}{units[idx[m]]}")
        print("  ".join(parts))


if __name__ == "__main__":
    main(sys.argv[1])
