#!/bin/bash
# pass 6 (ONE GPU): generalised CTA-pair halo kernel (C = 128 / 256) + deeper-ring CTA-pair gate kernel: parity, then A/B
set -u
TAG=${1:-r2_p6}
OUT=gpurun_out; mkdir -p $OUT
timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q -x 2>&1 | tail -4
CMTTS_GATE_PAIR=1 timeout 900 python -m pytest tests -m gpu -q -s > $OUT/gpu_tests_${TAG}.log 2>&1
tail -6 $OUT/gpu_tests_${TAG}.log; grep -E "mel max-abs|vocoder \[" $OUT/gpu_tests_${TAG}.log
run() { name=$1; shift; timeout 300 python bench.py --no-cpu-baseline --steps 20 --warmup 5 "$@" > $OUT/bench_${TAG}_$name.json 2> $OUT/bench_${TAG}_$name.err || tail -c 800 $OUT/bench_${TAG}_$name.err; }
CMTTS_HALO2=1 CMTTS_GATE_PAIR=0 run C2_h128_g1
CMTTS_HALO2=2 CMTTS_GATE_PAIR=0 run C2_h256_g1
CMTTS_HALO2=2 CMTTS_GATE_PAIR=1 run C2_h256_g2
CMTTS_HALO2=1 run C5_B32_h128 --config C5 --batch 32
CMTTS_HALO2=2 run C5_B32_h256 --config C5 --batch 32
python - <<PY
import json, glob
for f in sorted(glob.glob("$OUT/bench_${TAG}_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "unreadable", e); continue
    print(f.split("bench_${TAG}_")[1], round(d["ms_per_step"], 3), "ms", round(d["value"]), "fr/s", d["stages_ms"], "clk", d.get("clocks", {}).get("sm_mhz"))
    for k in d["kernels"][:12]:
        print("     ", round(k["ms"] / k["launches"] * 1e3, 1), "us x", k["launches"], k["kernel"])
PY
