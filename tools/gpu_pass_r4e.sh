#!/bin/bash
# New-component pass on ONE GPU: speaker-table + speaker-encoder GPU tests, C4 with the zero-shot speaker encoder in the step,
# C2 default line (sanity after the host-side changes), ncu launch list of the ResCNN.
set -u
TAG=${1:-r4e}
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi -L
timeout 600 python -m pytest tests/test_speaker_table.py tests/test_speaker_encoder.py tests/test_gpu_boundary.py -m gpu -q -s > $OUT/gpu_tests_$TAG.log 2>&1
tail -8 $OUT/gpu_tests_$TAG.log
run() { name=$1; shift; timeout 300 python bench.py --no-cpu-baseline --steps 20 --warmup 5 "$@" > $OUT/bench_${TAG}_$name.json 2> $OUT/bench_${TAG}_$name.err || tail -c 800 $OUT/bench_${TAG}_$name.err; }
run C4_zeroshot --config C4
run C4_plain --config C4 --zero-shot off
run C2_T4 --config C2
python - <<PY
import json, glob
for f in sorted(glob.glob("$OUT/bench_${TAG}_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "unreadable", e); continue
    print(f.split("bench_${TAG}_")[1], round(d["ms_per_step"], 3), "ms", round(d["value"]), "fr/s e2e", round(d["e2e"]["value"]), "h2d", d["e2e"]["h2d_bytes_per_step"],
          "stages", d.get("stages_ms"), "zs", d["config"].get("zero_shot_speaker_encoder"), "same_build", d["roofline"].get("traffic_same_build"))
    if "zeroshot" in f:
        for k in d.get("kernels", []):
            if "rescnn" in k["kernel"]:
                print("    ", k["kernel"], k["launches"], round(k["ms"], 4), "ms")
PY
cat > /tmp/rescnn_only.py <<PY
import sys, torch, numpy as np
sys.path.insert(0, ".")
from cmtts_b200 import speaker_encoder as SE, synthetic
m = SE.DeepSpeakerModel("cuda:0").set_keras_weights(synthetic.make_deepspeaker_weights(0))
x = torch.randn(1, 160, 64, device="cuda:0")
for _ in range(3):
    m.predict_tensor(x)
torch.cuda.synchronize()
PY
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/launches_${TAG}_rescnn.csv python /tmp/rescnn_only.py > $OUT/ncu_${TAG}_rescnn.log 2>&1
python tools/launch_summary.py $OUT/launches_${TAG}_rescnn.csv > $OUT/launches_${TAG}_rescnn_summary.txt; head -12 $OUT/launches_${TAG}_rescnn_summary.txt
