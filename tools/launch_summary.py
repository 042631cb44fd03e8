"""Summarise an ncu launch list (gpu__time_duration.sum CSV): time per kernel family and share of the step.

    python tools/launch_summary.py gpurun_out/launches.csv [--by-grid]
"""
import csv
import re
import sys
from collections import defaultdict


def short(name: str) -> str:
    name = re.sub(r"<unnamed>::", "", name)
    name = re.sub(r"^void\s+", "", name)
    m = re.match(r"([\w:]+)(<[^>]*>)?", name)
    return (m.group(1) + (m.group(2) or "")) if m else name[:60]


def main():
    path = sys.argv[1]
    by_grid = "--by-grid" in sys.argv
    rows = []
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        rows.append((short(r["Kernel Name"]), r["Grid Size"], r["Block Size"], float(r["Metric Value"]) / 1e3))
    # one step of the hot path = from one embed_tokens launch (first kernel of the encoder) to the next
    starts = [i for i, r in enumerate(rows) if r[0].startswith("embed_tokens")]
    if "--step" in sys.argv and len(starts) >= 2:
        k = int(sys.argv[sys.argv.index("--step") + 1])
        rows = rows[starts[k]:starts[k + 1]]
    tot = sum(r[3] for r in rows)
    agg = defaultdict(lambda: [0, 0.0])
    for n, g, b, t in rows:
        k = (n, g) if by_grid else (n,)
        agg[k][0] += 1
        agg[k][1] += t
    print(f"{len(rows)} launches, {tot / 1e3:.3f} ms total (ncu per-launch times are serialised and cold-cache)")
    print(f"{'kernel':70s} {'launches':>8s} {'us total':>10s} {'us avg':>8s} {'share':>6s}")
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{' '.join(k)[:70]:70s} {c:8d} {t:10.1f} {t / c:8.1f} {100 * t / tot:5.1f}%")


if __name__ == "__main__":
    main()
