#!/bin/bash
# Round-2 pass 1 on ONE GPU: full -m gpu suite, smoke, the default bench line + reference arm, every other BASELINE.json
# config (T=1, C1, C3, C4 weak/strong at N=1, C5 sweep), launch lists of a C2 and a C3 step, and ncu --set full of the
# vocoder's kernels at C2 size.
set -u
TAG=${1:-r2_p1}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi -L; nproc
timeout 900 python -m pytest tests -m gpu -x -q -s > $OUT/gpu_tests_$TAG.log 2>&1
tail -3 $OUT/gpu_tests_$TAG.log
timeout 120 python __graft_entry__.py --smoke 2>&1 | tail -2
timeout 400 python bench.py --steps 20 --warmup 5 > $OUT/bench_${TAG}_C2_T4.json 2> $OUT/bench_${TAG}_C2_T4.err
tail -c 3000 $OUT/bench_${TAG}_C2_T4.err
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_${TAG}_reference.json 2> $OUT/bench_${TAG}_reference.err
tail -c 3000 $OUT/bench_${TAG}_reference.err
run() { name=$1; shift; timeout 300 python bench.py --no-cpu-baseline --steps 20 --warmup 5 "$@" > $OUT/bench_${TAG}_$name.json 2> $OUT/bench_${TAG}_$name.err || tail -c 400 $OUT/bench_${TAG}_$name.err; }
run C2_T1 --config C2 --T 1
run C1 --config C1
run C3 --config C3
run C3_T4 --config C3 --T 4
run C4 --config C4
run C3_strong --config C3 --scaling strong
run C4_strong --config C4 --scaling strong
for B in 1 2 4 8 16 32 64 128 256; do run C5_B$B --config C5 --batch $B; done
python - <<PY
import json, glob
for f in sorted(glob.glob("$OUT/bench_${TAG}_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "unreadable", e); continue
    r = d.get("roofline", {})
    print(f.split("bench_${TAG}_")[1], round(d["ms_per_step"], 3), "ms", round(d["value"]), "fr/s e2e", round(d["e2e"]["value"]),
          "launches", d.get("gpu_launches"), "clk", d.get("clocks", {}).get("sm_mhz"), "| top", r.get("kernel"), round(r.get("frac", 0), 3),
          "share", round(r.get("share_of_step", 0), 3), "step_frac", round(r.get("step_frac", 0), 3))
PY
# launch lists (one step each)
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $OUT/launches_${TAG}_C2.csv \
    python tools/stage_only.py --stage all --config C2 --reps 2 > $OUT/ncu_${TAG}_C2.log 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $OUT/launches_${TAG}_C3.csv \
    python tools/stage_only.py --stage all --config C3 --reps 2 > $OUT/ncu_${TAG}_C3.log 2>&1
python tools/launch_summary.py $OUT/launches_${TAG}_C2.csv > $OUT/launches_${TAG}_C2_summary.txt; head -30 $OUT/launches_${TAG}_C2_summary.txt
python tools/launch_summary.py $OUT/launches_${TAG}_C3.csv > $OUT/launches_${TAG}_C3_summary.txt; head -30 $OUT/launches_${TAG}_C3_summary.txt
# full capture of the vocoder kernels at C2 size: second pass of the stage (skip the first = warm-up)
timeout 400 ncu --set full --clock-control none --import-source on -k regex:umma_ -s 64 -c 64 -o $OUT/prof_voc_$TAG -f \
    python tools/stage_only.py --stage vocoder --config C2 --reps 2 > $OUT/prof_voc_$TAG.log 2>&1
python tools/ncu_summary.py $OUT/prof_voc_$TAG.ncu-rep $OUT/ncu_${TAG}_voc_summary.csv
rm -f $OUT/prof_voc_$TAG.ncu-rep          # 120 MB: over gpurun's 64 MiB return limit; the per-launch summary is what is kept
du -sh $OUT; ls $OUT | wc -l
