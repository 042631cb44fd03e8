"""Per-launch summary of an ncu report (.ncu-rep): the metrics the roofline is argued from.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [out.csv]
"""
import csv
import io
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum" ,
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum",
    "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
    "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum", "sm__cycles_elapsed.max",
]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, body = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    have = [m for m in METRICS if m in col]
    out = [["kernel", "grid"] + have]
    for r in body:
        name = r[col["Kernel Name"]]
        out.append([name, r[col["Grid Size"]].replace(",", ";")] + [f"{r[col[m]]} {units[col[m]]}".strip() for m in have])
    w = csv.writer(open(sys.argv[2], "w", newline="") if len(sys.argv) > 2 else sys.stdout)
    w.writerows(out)


if __name__ == "__main__":
    main()
