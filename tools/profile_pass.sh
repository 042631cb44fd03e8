#!/bin/bash
# One profiling pass of the current build on a GPU box (about 4 minutes of box time):
#
#   gpurun --timeout 420 -- 'bash tools/profile_pass.sh TAG'
#
# writes into gpurun_out/: the -m gpu test log, one bench line (without the CPU baseline), the ncu launch list of one step
# with its per-kernel summary, `ncu --set full` captures of the vocoder's and the denoiser's tcgen05 kernels, and the
# per-role warp-state tables (tools/ncu_roles.py) of one launch of every distinct kernel in them.  Copy what is worth
# keeping into profiles/ (gpurun_out/ is scratch).
set -u
TAG=${1:-pass}
OUT=gpurun_out
mkdir -p $OUT

timeout 150 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > $OUT/gpu_tests_$TAG.log
tail -2 $OUT/gpu_tests_$TAG.log

timeout 120 python bench.py --no-cpu-baseline > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
python - <<EOF
import json
d = json.load(open("$OUT/bench_$TAG.json"))
print("bench", round(d["ms_per_step"], 2), "ms/step", d["stages_ms"], "SM MHz", d["clocks"]["sm_mhz"],
      "roofline frac", round(d["roofline"]["frac"], 3))
EOF

# launch list of one step (the first complete one: embed_tokens -> next embed_tokens)
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $OUT/ncu_bench_$TAG.log 2>&1
python tools/launch_summary.py $OUT/launches_$TAG.csv --step 0 > $OUT/launches_${TAG}_summary.txt 2>/dev/null
head -12 $OUT/launches_${TAG}_summary.txt

# full captures: vocoder (skip the first pass = warm-up), denoiser (skip the encoder / variance adaptor and one step)
timeout 200 ncu --set full --clock-control none --import-source on -k regex:umma_ -s 62 -c 62 -o $OUT/prof_voc_$TAG -f \
    python tools/vocoder_only.py > $OUT/prof_voc_$TAG.log 2>&1
timeout 120 ncu --set full --clock-control none --import-source on -k regex:umma_ -s 23 -c 8 -o $OUT/prof_dn_$TAG -f \
    python tools/denoiser_only.py > $OUT/prof_dn_$TAG.log 2>&1

for rep in voc dn; do
    f=$OUT/prof_${rep}_$TAG.ncu-rep
    [ -f $f ] || continue
    python tools/ncu_summary.py $f $OUT/ncu_${rep}_${TAG}_summary.csv > /dev/null 2>&1
    python - <<EOF > $OUT/ncu_${rep}_${TAG}_roles.txt 2>&1
import csv, re, subprocess, sys
rows = list(csv.reader(open("$OUT/ncu_${rep}_${TAG}_summary.csv")))
seen = {}
for i, r in enumerate(rows[1:]):
    m = re.search(r"(umma_\w+<[^>]*>)", r[0])
    k = m.group(1) if m else r[0][:40]
    if k not in seen:
        seen[k] = i
for k, i in seen.items():
    print("=" * 20, k, rows[1 + i][2])
    sys.stdout.flush()
    subprocess.run([sys.executable, "tools/ncu_roles.py", "$f", re.escape(k.replace(" ", ""))])
EOF
done
ls -la $OUT/*_$TAG*
