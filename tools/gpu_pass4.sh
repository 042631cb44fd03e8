#!/bin/bash
# pass 4 (ONE GPU): the CTA-pair halo kernel — op-level and vocoder parity with it on, then A/B timing
set -u
TAG=${1:-r2_p4}
OUT=gpurun_out; mkdir -p $OUT
timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q -x 2>&1 | tail -5
timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q -s -k "hifigan or vocoder or pipeline" > $OUT/gpu_tests_${TAG}_voc.log 2>&1
tail -6 $OUT/gpu_tests_${TAG}_voc.log; grep -E "vocoder \[" $OUT/gpu_tests_${TAG}_voc.log
run() { name=$1; shift; timeout 300 python bench.py --no-cpu-baseline --steps 20 --warmup 5 "$@" > $OUT/bench_${TAG}_$name.json 2> $OUT/bench_${TAG}_$name.err || tail -c 800 $OUT/bench_${TAG}_$name.err; }
CMTTS_HALO2=0 run C5_B32_halo1 --config C5 --batch 32
run C5_B32_halo2 --config C5 --batch 32
CMTTS_HALO2=0 run C2_halo1
run C2_halo2
python - <<PY
import json, glob
for f in sorted(glob.glob("$OUT/bench_${TAG}_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "unreadable", e); continue
    print(f.split("bench_${TAG}_")[1], round(d["ms_per_step"], 3), "ms", round(d["value"]), "fr/s", d["stages_ms"], "clk", d.get("clocks", {}).get("sm_mhz"))
    for k in d["kernels"][:10]:
        if "halo" in k["kernel"]: print("     ", round(k["ms"] / k["launches"] * 1e3, 1), "us x", k["launches"], k["kernel"])
PY
