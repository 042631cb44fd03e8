#!/bin/bash
# ncu --set full captures (source-level) of the layer GEMM / conditioner GEMM and of the paired-row ResBlock kernels.
set -u
TAG=${1:-r3c}
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi -L
# umma_conv_kernel launches of one T = 1 sampler pass: 24 in dpen, then the conditioner GEMM, the input projection, layer GEMMs
timeout 600 ncu --set full --clock-control none --import-source on -k regex:umma_conv_kernel -s 23 -c 6 -o $OUT/prof_rec_$TAG -f \
    python tools/stage_only.py --stage sampler --config C2 --T 1 --reps 1 > $OUT/prof_rec_$TAG.log 2>&1
tail -3 $OUT/prof_rec_$TAG.log
python tools/ncu_summary.py $OUT/prof_rec_$TAG.ncu-rep $OUT/ncu_${TAG}_rec_summary.csv; cat $OUT/ncu_${TAG}_rec_summary.csv | cut -c1-300
timeout 600 ncu --set full --clock-control none --import-source on -k regex:umma_resblock_kernel -s 6 -c 4 -o $OUT/prof_rbp_$TAG -f \
    python tools/stage_only.py --stage vocoder --config C2 --reps 1 > $OUT/prof_rbp_$TAG.log 2>&1
tail -3 $OUT/prof_rbp_$TAG.log
python tools/ncu_summary.py $OUT/prof_rbp_$TAG.ncu-rep $OUT/ncu_${TAG}_rbp_summary.csv; cat $OUT/ncu_${TAG}_rbp_summary.csv | cut -c1-300
ls -la $OUT/*.ncu-rep; du -sh $OUT
