#!/bin/bash
# 8 GPUs of one box, final build: BASELINE.json configs[2] (C3: VCTK, 64 utterances, T = 1) and configs[3] (C4: LibriTTS zero-shot,
# 128 utterances, T = 4) as they are stated — global batch over 8 GPUs — then the 8-rank equality check and the C2 line.
set -u
TAG=${1:-r4}; N=${2:-8}
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi -L | wc -l
for cfg in C3 C4 C2; do
  timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
      bench.py --gpus $N --steps 20 --warmup 5 --config $cfg --no-cpu-baseline > $OUT/bench_${TAG}_${cfg}_N${N}.json 2> $OUT/bench_${TAG}_${cfg}_N${N}.err
  tail -c 200 $OUT/bench_${TAG}_${cfg}_N${N}.err | grep -i "error" 
done
python - <<PY
import json, glob
for f in sorted(glob.glob("$OUT/bench_${TAG}_*_N${N}.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "unreadable", e); continue
    print(f.split("bench_${TAG}_")[1], "N", d.get("n_gpus"), round(d["ms_per_step"], 3), "ms", round(d["value"]), "fr/s e2e", round(d["e2e"]["value"]),
          "global batch", d["config"]["global_batch"], "pad/valid", round(d["config"]["padded_over_valid"], 4), "zs", d["config"].get("zero_shot_speaker_encoder"))
PY
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    tools/dist_check.py --full > $OUT/dist_check_${TAG}_${N}gpu.log 2>&1
grep -E "global|local|dist_check|Error|error" $OUT/dist_check_${TAG}_${N}gpu.log | tail -12
