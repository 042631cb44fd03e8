#!/bin/bash
# ncu --set full captures (source-level) of the gate conv and the layer GEMM inside one T = 1 sampler pass.
set -u
TAG=${1:-r3f}
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi -L
timeout 600 ncu --set full --clock-control none --import-source on -k regex:umma_gate8x2 -s 3 -c 2 -o $OUT/prof_gate_$TAG -f \
    python tools/stage_only.py --stage sampler --config C2 --T 1 --reps 1 > $OUT/prof_gate_$TAG.log 2>&1
tail -2 $OUT/prof_gate_$TAG.log
python tools/ncu_summary.py $OUT/prof_gate_$TAG.ncu-rep $OUT/ncu_${TAG}_gate_summary.csv; cat $OUT/ncu_${TAG}_gate_summary.csv | cut -c1-300
timeout 600 ncu --set full --clock-control none --import-source on -k regex:umma_conv_kernel -s 27 -c 2 -o $OUT/prof_rec_$TAG -f \
    python tools/stage_only.py --stage sampler --config C2 --T 1 --reps 1 > $OUT/prof_rec_$TAG.log 2>&1
tail -2 $OUT/prof_rec_$TAG.log
python tools/ncu_summary.py $OUT/prof_rec_$TAG.ncu-rep $OUT/ncu_${TAG}_rec_summary.csv; cat $OUT/ncu_${TAG}_rec_summary.csv | cut -c1-300
ls -la $OUT/*.ncu-rep; du -sh $OUT
