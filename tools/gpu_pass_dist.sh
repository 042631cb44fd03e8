#!/bin/bash
# N GPUs of one box: tools/dist_check.py --full (both padding modes, bench size) + the 2-rank pytest + one bench line per mode
set -u
TAG=${1:-r2_dist}; N=${2:-2}
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi -L | head -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    tools/dist_check.py --full > $OUT/dist_check_${TAG}_${N}gpu.log 2>&1
grep -E "global|local|dist_check|Error|error" $OUT/dist_check_${TAG}_${N}gpu.log | tail -20
timeout 300 python -m pytest tests/test_gpu_dist.py -m gpu -q 2>&1 | tail -3
for mode in balanced contiguous; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
      bench.py --gpus $N --steps 20 --warmup 5 --shard $mode > $OUT/bench_${TAG}_C2_N${N}_$mode.json 2> $OUT/bench_${TAG}_C2_N${N}_$mode.err
  tail -c 300 $OUT/bench_${TAG}_C2_N${N}_$mode.err
done
python - <<PY
import json, glob
for f in sorted(glob.glob("$OUT/bench_${TAG}_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "unreadable", e); continue
    print(f.split("bench_${TAG}_")[1], "N", d.get("n_gpus"), round(d["ms_per_step"], 3), "ms", round(d["value"]), "fr/s e2e", round(d["e2e"]["value"]),
          "pad/valid", round(d["config"]["padded_over_valid"], 4))
PY
