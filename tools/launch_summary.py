"""Per-kernel totals of an ncu launch list (`ncu --metrics gpu__time_duration.sum --csv --log-file X.csv ...`):

    python tools/launch_summary.py gpurun_out/launches.csv [--skip N]

Prints kernel, launches, total us, share — cold-cache, serialised times: compare SHARES, not absolutes."""
import csv
import re
import sys
from collections import OrderedDict


def short(name):
    m = re.search(r"(\w+)<([^(]*)>\(", name)
    if m:
        return f"{m.group(1)}<{m.group(2)}>"
    return name.split("(")[0][-70:]


def main():
    path = sys.argv[1]
    skip = int(sys.argv[sys.argv.index("--skip") + 1]) if "--skip" in sys.argv else 0
    lines = [l for l in open(path, errors="replace") if l.startswith('"')]
    rows = list(csv.DictReader(lines))
    agg = OrderedDict()
    n = 0
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        n += 1
        if n <= skip:
            continue
        v = float(r["Metric Value"].replace(",", ""))
        u = r["Metric Unit"].lower()
        us = v * {"nsecond": 1e-3, "ns": 1e-3, "usecond": 1.0, "us": 1.0, "msecond": 1e3, "ms": 1e3}.get(u, 1e-3)
        k = short(r["Kernel Name"])
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += us
    tot = sum(a[1] for a in agg.values()) or 1.0
    print(f"{n - skip} launches, {tot / 1e3:.3f} ms total")
    for k, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{us:10.1f} us  {100 * us / tot:5.1f} %  x{c:<4d} {us / c:8.1f} us/launch  {k}")


if __name__ == "__main__":
    main()
