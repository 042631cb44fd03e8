"""Timing ablations of the tcgen05 kernels (run on the GPU box): which role bounds a kernel?

Runs tools/vocoder_only.py / tools/denoiser_only.py under an ncu launch list once per debug setting
(CMTTS_RB_DBG / CMTTS_UMMA_DBG / CMTTS_PF: roles switched off one at a time — results are wrong, only the
durations matter) and prints the average duration of every kernel family per setting.
"""
import csv
import os
import re
import subprocess
import sys
from collections import OrderedDict, defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(script, env_extra, tag):
    out = os.path.join(ROOT, "gpurun_out", f"abl_{tag}.csv")
    env = dict(os.environ, **env_extra)
    cmd = ["ncu", "--metrics", "gpu__time_duration.sum", "--clock-control", "none", "-c", "400", "--csv", "--log-file", out,
           sys.executable, os.path.join(ROOT, "tools", script)]
    subprocess.run(cmd, env=env, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, timeout=240)
    agg = defaultdict(lambda: [0, 0.0])
    with open(out) as f:
        lines = [l for l in f if l.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"<unnamed>::|^void\s+", "", r["Kernel Name"])
        m = re.match(r"([\w:]+)(<[^>]*>)?", name)
        k = (m.group(1) + (m.group(2) or "")) if m else name[:40]
        if "umma" not in k:
            continue
        agg[k][0] += 1
        agg[k][1] += float(r["Metric Value"]) / 1e3
    os.remove(out)
    return {k: v[1] / v[0] for k, v in agg.items()}


def table(title, results):
    keys = sorted({k for r in results.values() for k in r})
    print(f"== {title}: average us per launch")
    print(f"{'kernel':44s}" + "".join(f"{t:>12s}" for t in results))
    for k in keys:
        print(f"{k[:44]:44s}" + "".join(f"{r.get(k, float('nan')):12.1f}" for r in results.values()))
    sys.stdout.flush()


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which in ("all", "voc"):
        res = OrderedDict()
        for tag, env in [("base", {}), ("noE2", {"CMTTS_RB_DBG": "1"}), ("noE1", {"CMTTS_RB_DBG": "2"}),
                         ("noMMA", {"CMTTS_RB_DBG": "4"}), ("noE12MMA", {"CMTTS_RB_DBG": "7"}),
                         ("noStore", {"CMTTS_RB_DBG": "16"}), ("pf8", {"CMTTS_PF": "8"})]:
            res[tag] = run("vocoder_only.py", env, "voc_" + tag)
        table("vocoder", res)
    if which == "ring":
        res = OrderedDict()
        for tag, env in [("base", {}), ("nb2", {"CMTTS_RB_NB": "2"}), ("skel", {"CMTTS_RB_DBG": "7"}),
                         ("skel+pf8", {"CMTTS_RB_DBG": "7", "CMTTS_PF": "8"}), ("skel+nb2", {"CMTTS_RB_DBG": "7", "CMTTS_RB_NB": "2"}),
                         ("noMMA+nb2", {"CMTTS_RB_DBG": "4", "CMTTS_RB_NB": "2"})]:
            res[tag] = run("vocoder_only.py", env, "ring_" + tag)
        table("vocoder input-ring experiments", res)
    if which == "tma":
        res = OrderedDict()
        for tag, env in [("skel", {"CMTTS_RB_DBG": "7"}), ("skel16rows", {"CMTTS_RB_DBG": "39"}),
                         ("skelNoRes", {"CMTTS_RB_DBG": "15"}), ("skel16NoRes", {"CMTTS_RB_DBG": "47"})]:
            res[tag] = run("vocoder_only.py", env, "tma_" + tag)
        table("TMA row-rate experiment (fused ResBlock kernel, all roles but the producer switched off)", res)
    if which in ("all", "dn"):
        res = OrderedDict()
        for tag, env in [("base", {}), ("noStore", {"CMTTS_UMMA_DBG": "16"}), ("noMMA", {"CMTTS_UMMA_DBG": "32"}),
                         ("noEpi", {"CMTTS_UMMA_DBG": "64"}), ("noMMAnoEpi", {"CMTTS_UMMA_DBG": "96"})]:
            res[tag] = run("denoiser_only.py", env, "dn_" + tag)
        table("denoiser", res)


if __name__ == "__main__":
    main()
