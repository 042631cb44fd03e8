"""Top stall sites of one profiled launch (ncu --page source --csv), SASS view.

    python tools/ncu_stalls.py report.ncu-rep <launch index> [top N]
"""
import csv
import io
import subprocess
import sys


def main():
    rep, idx = sys.argv[1], sys.argv[2]
    top_n = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = [r for r in csv.reader(io.StringIO(raw))]
    starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"] + [len(rows)]
    k = int(idx)
    rows = rows[starts[k]:starts[k + 1]]
    print(rows[0][1][:120])
    hdr = rows[1]
    body = [r for r in rows[2:] if len(r) == len(hdr) and r[0].startswith("0x")]
    ci = {h: i for i, h in enumerate(hdr)}
    num = lambda r, c: int(float(r[ci[c]] or 0))
    tot = sum(num(r, "# Samples") for r in body)
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    print("total samples", tot)
    for r in sorted(body, key=lambda r: -num(r, "# Samples"))[:top_n]:
        s = num(r, "# Samples")
        st = sorted(((num(r, c), c[6:]) for c in stall_cols), reverse=True)[:2]
        print(f"{s:6d} {100 * s / max(tot, 1):5.1f}%  {r[ci['Address']][-5:]} {r[ci['Source']][:80]:80s} {st}")


if __name__ == "__main__":
    main()
