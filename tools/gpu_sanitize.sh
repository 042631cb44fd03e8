#!/bin/bash
# compute-sanitizer over the speaker encoder (memcheck + racecheck: shared-memory reduction across warps) and a small pass of
# the whole hot path (memcheck on the smoke-size pipeline)
set -u
TAG=${1:-r4}
OUT=gpurun_out; mkdir -p $OUT
cat > /tmp/san_rescnn.py <<PY
import sys, torch, numpy as np
sys.path.insert(0, ".")
from cmtts_b200 import speaker_encoder as SE, synthetic
m = SE.DeepSpeakerModel("cuda:0").set_keras_weights(synthetic.make_deepspeaker_weights(0))
for B, T in ((2, 160), (1, 37)):
    x = torch.randn(B, T, 64)
    e = m.predict_tensor(x)
    torch.cuda.synchronize()
    print(B, T, float(e.norm(dim=1).mean()))
PY
for tool in memcheck racecheck; do
  timeout 600 compute-sanitizer --tool $tool --print-limit 5 python /tmp/san_rescnn.py > $OUT/sanitizer_${TAG}_rescnn_$tool.log 2>&1
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|^[0-9] " $OUT/sanitizer_${TAG}_rescnn_$tool.log | tail -4
done
timeout 900 compute-sanitizer --tool memcheck --print-limit 5 python __graft_entry__.py --smoke > $OUT/sanitizer_${TAG}_smoke_memcheck.log 2>&1
grep -E "ERROR SUMMARY|smoke ok" $OUT/sanitizer_${TAG}_smoke_memcheck.log | tail -3
