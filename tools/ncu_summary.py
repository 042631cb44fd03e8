"""Per-launch summary of an `ncu --set full` report (read here, no GPU needed):

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [out.csv]

One row per captured launch: kernel, duration, tensor-pipe activity, DRAM bytes / throughput, L2->SM bytes, achieved
occupancy, registers.  Also usable on launch lists (`--metrics gpu__time_duration.sum` CSV logs) via tools/launch_summary.py.
"""
import csv
import io
import re
import subprocess
import sys

METRICS = [
    ("gpu__time_duration.sum", "us", 1e-3),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pipe_pct", 1.0),
    ("sm__inst_executed_pipe_tensor.sum", "tensor_insts", 1.0),
    ("dram__bytes_read.sum", "dram_read_MB", 1e-6),
    ("dram__bytes_write.sum", "dram_write_MB", 1e-6),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct", 1.0),
    ("lts__t_bytes.sum", "l2_bytes_MB", 1e-6),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy_pct", 1.0),
    ("launch__registers_per_thread", "regs", 1.0),
    ("launch__grid_size", "grid", 1.0),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct", 1.0),
]


def short(name):
    m = re.search(r"(\w+)<([^(]*)>\(", name)
    if m:
        return f"{m.group(1)}<{m.group(2)}>"
    return name.split("(")[0][-60:]


def to_bytes(v, unit):
    u = unit.lower()
    f = {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "tbyte": 1e12}.get(u)
    return v * f if f else v


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    out = [["kernel"] + [m[1] for m in METRICS]]
    for r in data:
        line = [short(r[col["Kernel Name"]])]
        for name, _, scale in METRICS:
            if name not in col:
                line.append("")
                continue
            try:
                v = float(r[col[name]].replace(",", ""))
            except ValueError:
                line.append(r[col[name]])
                continue
            unit = units[col[name]]
            if "byte" in unit.lower():
                v = to_bytes(v, unit) * scale
            elif name == "gpu__time_duration.sum":
                v = v * {"nsecond": 1e-3, "ns": 1e-3, "usecond": 1.0, "us": 1.0, "msecond": 1e3, "ms": 1e3, "second": 1e6, "s": 1e6}.get(unit.lower(), 1e-3)
            else:
                v = v * scale
            line.append(f"{v:.3f}")
        out.append(line)
    w = csv.writer(open(sys.argv[2], "w", newline="") if len(sys.argv) > 2 else sys.stdout)
    w.writerows(out)


if __name__ == "__main__":
    main()
