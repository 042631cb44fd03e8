#!/bin/bash
# Quick pass on ONE GPU: the -m gpu suite, smoke, the C2 bench line (T = 4 and 1), C3, and the launch list of one C2 step.
set -u
TAG=${1:-quick}
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi -L
timeout 1200 python -m pytest tests -m gpu -q -s > $OUT/gpu_tests_$TAG.log 2>&1
grep -E 'paired-row|FAILED|Error|error' $OUT/gpu_tests_$TAG.log | head -20; tail -6 $OUT/gpu_tests_$TAG.log
timeout 120 python __graft_entry__.py --smoke 2>&1 | tail -1
run() { name=$1; shift; timeout 300 python bench.py --no-cpu-baseline --steps 20 --warmup 5 "$@" > $OUT/bench_${TAG}_$name.json 2> $OUT/bench_${TAG}_$name.err || tail -c 600 $OUT/bench_${TAG}_$name.err; }
run C2_T4 --config C2
run C2_T1 --config C2 --T 1
run C3 --config C3
run C5_B32 --config C5 --batch 32
CMTTS_RB_PAIR=0 run C5_B32_nopair --config C5 --batch 32
python - <<PY
import json, glob
for f in sorted(glob.glob("$OUT/bench_${TAG}_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "unreadable", e); continue
    r = d.get("roofline", {})
    print(f.split("bench_${TAG}_")[1], round(d["ms_per_step"], 3), "ms", round(d["value"]), "fr/s e2e", round(d["e2e"]["value"]),
          "launches", d.get("gpu_launches"), "clk", d.get("clocks", {}).get("sm_mhz"), "stages", d.get("stages_ms"), "| top", r.get("kernel"), round(r.get("frac", 0), 3),
          "step_frac", round(r.get("step_frac", 0), 3))
    if "C2_T4" in f:
        for k in d.get("kernels", [])[:14]:
            print("    ", k["kernel"], k["launches"], round(k["ms"], 3), "ms")
PY
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $OUT/launches_${TAG}_C2.csv \
    python tools/stage_only.py --stage all --config C2 --reps 2 > $OUT/ncu_${TAG}_C2.log 2>&1
python tools/launch_summary.py $OUT/launches_${TAG}_C2.csv > $OUT/launches_${TAG}_C2_summary.txt; head -30 $OUT/launches_${TAG}_C2_summary.txt
du -sh $OUT
