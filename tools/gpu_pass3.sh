#!/bin/bash
# pass 3 (ONE GPU): full -m gpu suite, smoke, C1/C3 with and without CUDA graphs, default line
set -u
TAG=${1:-r2_p3}
OUT=gpurun_out
mkdir -p $OUT
CMTTS_GATE_FP8=0 timeout 1200 python -m pytest tests -m gpu -q -s > $OUT/gpu_tests_${TAG}_fp16x.log 2>&1
tail -12 $OUT/gpu_tests_${TAG}_fp16x.log
grep -E "mel max-abs|vocoder \[|RNG stream|Warning|warn" $OUT/gpu_tests_${TAG}_fp16x.log | head -20
echo "=== e4m3 cross terms in the gate conv (default) ==="
timeout 1200 python -m pytest tests -m gpu -q -s > $OUT/gpu_tests_$TAG.log 2>&1
tail -12 $OUT/gpu_tests_$TAG.log
grep -E "mel max-abs|vocoder \[|RNG stream|Warning|warn" $OUT/gpu_tests_$TAG.log | head -20
timeout 120 python __graft_entry__.py --smoke 2>&1 | tail -2
run() { name=$1; shift; timeout 300 python bench.py --no-cpu-baseline --steps 30 --warmup 5 "$@" > $OUT/bench_${TAG}_$name.json 2> $OUT/bench_${TAG}_$name.err || tail -c 800 $OUT/bench_${TAG}_$name.err; grep -i "warn" $OUT/bench_${TAG}_$name.err | head -3; }
run C1_graphs --config C1 --graphs on
run C1_eager --config C1 --graphs off
run C3_graphs --config C3 --graphs on
run C3_eager --config C3 --graphs off
run C3T4_graphs --config C3 --T 4 --graphs on
run C2T1_graphs --config C2 --T 1 --graphs on
run C2T1_eager --config C2 --T 1 --graphs off
CMTTS_GATE_FP8=0 timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $OUT/bench_${TAG}_C2_T4_fp16x.json 2> $OUT/bench_${TAG}_C2_T4_fp16x.err
timeout 400 python bench.py --steps 20 --warmup 5 > $OUT/bench_${TAG}_C2_T4.json 2> $OUT/bench_${TAG}_C2_T4.err
tail -c 300 $OUT/bench_${TAG}_C2_T4.err
python - <<PY
import json, glob
for f in sorted(glob.glob("$OUT/bench_${TAG}_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "unreadable", e); continue
    print(f.split("bench_${TAG}_")[1], round(d["ms_per_step"], 3), "ms", round(d["value"]), "fr/s e2e", round(d["e2e"]["value"]), round(d["e2e"]["ms_per_step"], 3),
          "launches", d.get("gpu_launches"), "graphs", d["config"].get("cuda_graphs"), d["config"].get("graph_replays"), "clk", d.get("clocks", {}).get("sm_mhz"))
PY
du -sh $OUT
