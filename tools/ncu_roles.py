"""Per-ROLE warp-state samples of a warp-specialised kernel from an ncu report (needs --set full --import-source on).

A tcgen05 kernel here is an if / else chain over the warp index: TMA producer(s), MMA issuer(s), epilogue warps.  ncu's
source page has per-SASS-instruction sample counts; this tool cuts the address space of one launch into the code ranges of
the roles — recognised by their landmark instructions (UTMALDG / UTCHMMA / LDTM), ranges split at the last backward branch
(= end of a role's tile loop) between two landmarks of different kinds — and prints, per role: samples, samples spent
spinning on an mbarrier (try_wait loop), top stall reasons and top opcodes.  A role that never spins is the bottleneck.

    python tools/ncu_roles.py report.ncu-rep [launch index | kernel-name regex, default 0]

(the source page may list a launch more than once, so its indices are not those of the raw page: select by name)
"""
import collections
import csv
import io
import re
import subprocess
import sys

LANDMARKS = (("UTMALDG", "TMA producer"), ("UTCHMMA", "MMA issuer"), ("LDTM", "epilogue"))


def load(rep, idx):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"] + [len(rows)]
    if not isinstance(idx, int):
        hits = [k for k, i in enumerate(starts[:-1]) if re.search(idx, rows[i][1].replace("(int)", "").replace(" ", ""))]
        if not hits:
            raise SystemExit(f"no launch matches {idx!r}")
        idx = hits[0]
    rows = rows[starts[idx]:starts[idx + 1]]
    name, hdr = rows[0][1], rows[1]
    body = [r for r in rows[2:] if len(r) == len(hdr) and r[0].startswith("0x")]
    return name, hdr, body


def main():
    rep = sys.argv[1]
    idx = sys.argv[2] if len(sys.argv) > 2 else "0"
    idx = int(idx) if idx.isdigit() else idx
    name, hdr, body = load(rep, idx)
    ci = {h: i for i, h in enumerate(hdr)}
    addr = [int(r[ci["Address"]], 16) for r in body]
    src = [r[ci["Source"]].strip() for r in body]
    smp = [int(float(r[ci["# Samples"]] or 0)) for r in body]
    exe = [int(float(r[ci["Instructions Executed"]] or 0)) for r in body]
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]

    def kind(s):
        for key, role in LANDMARKS:
            if re.search(r"\b" + key, s):
                return role
        return None

    marks = [(i, kind(s)) for i, s in enumerate(src) if kind(s) and exe[i] > 0]
    if not marks:
        print("no executed landmark instructions found")
        return
    # segment boundaries: between consecutive landmarks of different kinds, cut after the last backward branch
    cuts = [0]
    labels = [marks[0][1]]
    for (i0, k0), (i1, k1) in zip(marks, marks[1:]):
        if k0 == k1:
            continue
        cut = None
        for j in range(i1, i0, -1):
            m = re.search(r"\bBRA(?:\.\w+)*\s+(?:!?U?P\d+,\s*)?(0x[0-9a-f]+)", src[j])
            if m and int(m.group(1), 16) <= addr[j]:
                cut = j + 1
                break
        cuts.append(cut if cut is not None else (i0 + i1) // 2)
        labels.append(k1)
    # the code after the last role's loop (final barrier, TMEM dealloc, cold paths)
    last = marks[-1][0]
    tail = None
    for j in range(last, len(src)):
        if "BAR.SYNC" in src[j]:
            tail = j
            break
    cuts.append(tail if tail is not None else len(src))
    if tail is not None:
        cuts.append(len(src))
        labels.append("tail (final barrier, cold paths)")

    # spin loops: a short backward branch whose body contains an mbarrier try_wait (SYNCS...TRYWAIT); every sample inside
    # the body counts as "waiting for another role".  Cold-path copies of the loop (placed after the kernel's EXIT by
    # ptxas) branch FORWARD out of the loop, so a body is also recognised by "TRYWAIT ... @!P BRA <own address - small>".
    where = {a: i for i, a in enumerate(addr)}
    spin = [False] * len(src)
    for j, s_ in enumerate(src):
        m = re.search(r"\bBRA(?:\.\w+)*\s+(?:!?U?P\d+,\s*)?(0x[0-9a-f]+)", s_)
        if not m:
            continue
        tgt = where.get(int(m.group(1), 16))
        if tgt is None or tgt > j or j - tgt > 8:
            continue
        if any("TRYWAIT" in src[k] for k in range(tgt, j + 1)):
            for k in range(tgt, j + 1):
                spin[k] = True
    # the hot copy of a wait is "TRYWAIT ; ... ; @!P BRA cold_loop": the branch instruction collects the samples
    for j, s_ in enumerate(src):
        if "BRA" in s_ and s_.startswith("@") and any("TRYWAIT" in src[k] for k in range(max(0, j - 6), j)):
            spin[j] = True

    total = sum(smp)
    print(name[:140])
    print(f"{total} samples over {len(body)} SASS instructions")
    print(f"{'role':34s} {'addr range':>15s} {'samples':>8s} {'spin':>6s}  top stalls | top opcodes")
    seen = collections.Counter()
    for lo, hi, lab in zip(cuts, cuts[1:], labels):
        seen[lab] += 1
        t = w = 0
        st, ops = collections.Counter(), collections.Counter()
        for j in range(lo, hi):
            t += smp[j]
            if spin[j]:
                w += smp[j]
            m = re.match(r"(@!?U?P\d\s+)?(\S+)", src[j])
            ops[m.group(2).split(".")[0]] += smp[j]
            for c in stall_cols:
                v = int(float(body[j][ci[c]] or 0))
                if v:
                    st[c[6:]] += v
        tag = lab if seen[lab] == 1 else f"{lab} #{seen[lab]}"
        rng = f"{addr[lo] & 0xfffff:05x}-{addr[hi - 1] & 0xfffff:05x}"
        print(f"{tag:34s} {rng:>15s} {t:8d} {w:6d}  {dict(st.most_common(4))} | {ops.most_common(5)}")


if __name__ == "__main__":
    main()
