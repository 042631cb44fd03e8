#!/bin/bash
# Round-2 pass 2 on TWO GPUs of one box: the whole -m gpu suite (incl. the 2-rank NCCL test), tools/dist_check.py at bench
# size, the reference arm, and 2-GPU bench lines in both sharding modes.
set -u
TAG=${1:-r2_p2}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi -L; nproc
timeout 1200 python -m pytest tests -m gpu -q -s > $OUT/gpu_tests_$TAG.log 2>&1
tail -15 $OUT/gpu_tests_$TAG.log
grep -E "mel max-abs|float errors|vocoder:" $OUT/gpu_tests_$TAG.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    tools/dist_check.py --full > $OUT/dist_check_${TAG}_2gpu.log 2>&1
grep -E "global|local|dist_check|Error|error" $OUT/dist_check_${TAG}_2gpu.log | tail -20
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_${TAG}_reference.json 2> $OUT/bench_${TAG}_reference.err
tail -c 1500 $OUT/bench_${TAG}_reference.err; cut -c1-400 $OUT/bench_${TAG}_reference.json
for mode in balanced contiguous; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
      bench.py --gpus 2 --steps 20 --warmup 5 --shard $mode > $OUT/bench_${TAG}_C2_N2_$mode.json 2> $OUT/bench_${TAG}_C2_N2_$mode.err
  tail -c 600 $OUT/bench_${TAG}_C2_N2_$mode.err
done
timeout 300 python bench.py --steps 20 --warmup 5 > $OUT/bench_${TAG}_C2_N1.json 2> $OUT/bench_${TAG}_C2_N1.err
python - <<PY
import json, glob
for f in sorted(glob.glob("$OUT/bench_${TAG}_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "unreadable", e); continue
    print(f.split("bench_${TAG}_")[1], "N", d.get("n_gpus"), round(d["ms_per_step"], 3), "ms", round(d["value"]), "fr/s e2e", round(d["e2e"]["value"]),
          "pad/valid", d.get("config", {}).get("padded_over_valid"), "cpu", d.get("cpu_baseline", {}).get("kind"), d.get("cpu_baseline", {}).get("value"))
PY
du -sh $OUT
