#!/bin/bash
# pass 5 (ONE GPU): CTA-pair gate kernel — parity with it on, then A/B timing
set -u
TAG=${1:-r2_p5}
OUT=gpurun_out; mkdir -p $OUT
CMTTS_GATE_PAIR=1 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q -s -x > $OUT/gpu_tests_${TAG}_pair.log 2>&1
tail -6 $OUT/gpu_tests_${TAG}_pair.log; grep -E "mel max-abs" $OUT/gpu_tests_${TAG}_pair.log
run() { name=$1; shift; timeout 300 python bench.py --no-cpu-baseline --steps 20 --warmup 5 "$@" > $OUT/bench_${TAG}_$name.json 2> $OUT/bench_${TAG}_$name.err || tail -c 800 $OUT/bench_${TAG}_$name.err; }
CMTTS_GATE_PAIR=0 run C2_gate1
CMTTS_GATE_PAIR=1 run C2_gate2
CMTTS_GATE_PAIR=0 run C2_gate1b
CMTTS_GATE_PAIR=1 run C2_gate2b
python - <<PY
import json, glob
for f in sorted(glob.glob("$OUT/bench_${TAG}_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "unreadable", e); continue
    print(f.split("bench_${TAG}_")[1], round(d["ms_per_step"], 3), "ms", round(d["value"]), "fr/s", d["stages_ms"], "clk", d.get("clocks", {}).get("sm_mhz"))
    for k in d["kernels"][:6]:
        if "gate" in k["kernel"] or "e5" in k["kernel"]: print("     ", round(k["ms"] / k["launches"] * 1e3, 1), "us x", k["launches"], k["kernel"])
PY
