"""profiles/roofline_traffic_r2.json from ncu --set full summaries (tools/ncu_summary.py CSVs):

    python tools/make_traffic.py profiles/ncu_r2_*_summary.csv

Per kernel (template instantiation) the mean DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum), keyed by
the PREFIX of the launch profiler's label (bench.py looks a label up by its longest matching key), plus the stamp of the
library build the captures were taken with (cmtts_b200/lib/libcmtts_b200.so.stamp)."""
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EPI = {"0": "e0", "1": "e1", "2": "e2", "3": "e3", "4": "e4", "5": "e5"}


def label_prefix(kernel: str):
    m = re.match(r"umma_gate8x2_kernel<(\d+)>", kernel)
    if m:
        return f"umma_gate8x2<{m.group(1)}>"
    m = re.match(r"umma_gate8_kernel<(\d+)>", kernel)
    if m:
        return f"umma_gate8<{m.group(1)}>"
    m = re.match(r"umma_gate_kernel<(\d+), *\d+>", kernel)
    if m:
        return f"umma_gate<{m.group(1)}>"
    m = re.match(r"umma_halo2_kernel<(\d+), *(\d+)>", kernel)
    if m:
        return f"umma_halo2<{m.group(1)}> k{m.group(2)} "
    m = re.match(r"umma_halo_kernel<(\d+), *\d+, *\d+, *(\d+)>", kernel)
    if m:
        return f"umma_halo<{m.group(1)}> k{m.group(2)} "
    m = re.match(r"umma_resblock_kernel<(\d+), *(\d+), *\d+>", kernel)
    if m:
        return f"umma_resblock<{m.group(1)}> k{m.group(2)} "
    m = re.match(r"umma_conv_kernel<(\d+), *(\d+), *(\d+), *\d+, *(\d+)>", kernel)
    if m:
        return f"umma_conv<{m.group(1)},{m.group(2)},{m.group(3)},e{m.group(4)}>"
    for pat, lab in (("conv_post_f16", "conv_post_f16"), ("length_regulate", "length_regulate"), ("f32_to_f16_kernel", "f32_to_f16"),
                     ("renoise", "renoise")):
        if pat in kernel:
            return lab
    return None


def main():
    acc = {}
    for path in sys.argv[1:]:
        rows = list(csv.reader(open(path)))
        col = {h: i for i, h in enumerate(rows[0])}
        for r in rows[1:]:
            key = label_prefix(r[0])
            if key is None:
                continue
            try:
                b = (float(r[col["dram_read_MB"]]) + float(r[col["dram_write_MB"]])) * 1e6
                t = float(r[col["us"]])
            except (ValueError, KeyError):
                continue
            a = acc.setdefault(key, {"n": 0, "bytes": 0.0, "us": 0.0, "pipe": 0.0, "source": os.path.basename(path)})
            a["n"] += 1; a["bytes"] += b; a["us"] += t
            try:
                a["pipe"] += float(r[col["tensor_pipe_pct"]])
            except (ValueError, KeyError):
                pass
    stamp = None
    sp = os.path.join(ROOT, "cmtts_b200", "lib", "libcmtts_b200.so.stamp")
    if os.path.isfile(sp):
        stamp = open(sp).read().strip()
    sys.path.insert(0, ROOT)
    from cmtts_b200.build import kernel_stamp
    out = {"_build_stamp": stamp, "_kernel_stamp": kernel_stamp(), "_note": "mean per launch over the captured launches of each kernel; ncu --set full --clock-control none"}
    for k, a in sorted(acc.items()):
        out[k] = {"dram_bytes_per_launch": a["bytes"] / a["n"], "launches_captured": a["n"], "us_per_launch_under_ncu": a["us"] / a["n"],
                  "tensor_pipe_pct": a["pipe"] / a["n"], "source": a["source"]}
    json.dump(out, open(os.path.join(ROOT, "profiles", "roofline_traffic_r2.json"), "w"), indent=1)
    print(json.dumps(out, indent=1)[:3000])


if __name__ == "__main__":
    main()
