"""Reduce an .ncu-rep to a small CSV (one row per profiled launch) with the metrics DESIGN.md / profiles/README.md quote.
Runs on the GPU box right after the capture so only the CSV has to travel back:
    python tools/ncu_extract.py gpurun_out/prof.ncu-rep profiles_out.csv"""
import csv
import subprocess
import sys

WANT = ["ID", "Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "smsp__inst_executed.sum", "sm__cycles_elapsed.avg", "smsp__cycles_active.avg"]


def main() -> None:
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = [hdr.index(w) for w in WANT if w in hdr]
    # tensor-pipe columns have arch-specific names: keep anything that mentions the tensor pipe
    idx += [i for i, h in enumerate(hdr) if "pipe_tensor" in h and i not in idx][:6]
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow([hdr[i] for i in idx])
        w.writerow([units[i] for i in idx])
        for r in data:
            w.writerow([r[i] for i in idx])
    print(f"{out}: {len(data)} launches, {len(idx)} columns")


if __name__ == "__main__":
    main()
