"""One line of hardware counters per kernel of `ncu --set full` captures (gpurun_out/*.ncu-rep) -> profiles/ text.
usage: summarize_ncu_full.py out.txt rep1.ncu-rep [rep2 ...]   (needs the `ncu` CLI: build container, no GPU)"""
import csv
import io
import json
import subprocess
import sys

KEYS = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pipe_pct"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_pct"),
        ("launch__registers_per_thread", "regs"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
        ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("launch__shared_mem_per_block_dynamic", "dyn_smem")]


def rows_of(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        d = {"kernel": r[idx["Kernel Name"]]}
        for k, short in KEYS:
            if k in idx:
                d[short] = f"{r[idx[k]]} {units[idx[k]]}".strip()
        yield d


def main():
    dst, reps = sys.argv[1], sys.argv[2:]
    lines, js = [], []
    for rep in reps:
        lines.append(f"== {rep}")
        for d in rows_of(rep):
            lines.append("  " + d.pop("kernel"))
            lines.append("    " + "  ".join(f"{k}={v}" for k, v in d.items()))
            js.append(d)
    open(dst, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()
