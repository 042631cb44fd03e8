#!/bin/bash
# speaker-encoder re-check after the kernel rewrite: GPU tests, C4 with the zero-shot encoder in the step, ResCNN launch list
set -u
TAG=${1:-r4f}
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi -L
timeout 600 python -m pytest tests/test_speaker_encoder.py -m gpu -q -s > $OUT/gpu_tests_$TAG.log 2>&1
tail -7 $OUT/gpu_tests_$TAG.log
timeout 300 python bench.py --no-cpu-baseline --steps 20 --warmup 5 --config C4 > $OUT/bench_${TAG}_C4_zeroshot.json 2> $OUT/bench_${TAG}_C4_zeroshot.err || tail -c 800 $OUT/bench_${TAG}_C4_zeroshot.err
python - <<PY
import json
d = json.loads(open("$OUT/bench_${TAG}_C4_zeroshot.json").read().strip().splitlines()[-1])
print(round(d["ms_per_step"], 3), "ms", round(d["value"]), "fr/s e2e", round(d["e2e"]["value"]), "stages", d.get("stages_ms"), d["config"].get("zero_shot_speaker_encoder"))
for k in d.get("kernels", []):
    if "rescnn" in k["kernel"]:
        print("    ", k["kernel"], k["launches"], round(k["ms"], 4), "ms")
PY
cat > /tmp/rescnn_only.py <<PY
import sys, torch
sys.path.insert(0, ".")
from cmtts_b200 import speaker_encoder as SE, synthetic
m = SE.DeepSpeakerModel("cuda:0").set_keras_weights(synthetic.make_deepspeaker_weights(0))
x = torch.randn(1, 160, 64, device="cuda:0")
for _ in range(3):
    m.predict_tensor(x)
torch.cuda.synchronize()
PY
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/launches_${TAG}_rescnn.csv python /tmp/rescnn_only.py > $OUT/ncu_${TAG}_rescnn.log 2>&1
python tools/launch_summary.py $OUT/launches_${TAG}_rescnn.csv > $OUT/launches_${TAG}_rescnn_summary.txt; head -10 $OUT/launches_${TAG}_rescnn_summary.txt
