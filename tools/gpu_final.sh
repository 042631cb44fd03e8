#!/bin/bash
# Final pass of the round on ONE GPU: the whole -m gpu suite, smoke, every BASELINE.json config as a bench line, launch lists,
# and ncu --set full summaries (denoiser kernels, vocoder kernels, row kernels).  Everything lands in gpurun_out/ (< 64 MiB:
# the .ncu-rep files are summarised on the box and deleted).
set -u
TAG=${1:-r2_final}
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi -L; nproc
timeout 1200 python -m pytest tests -m gpu -q -s > $OUT/gpu_tests_$TAG.log 2>&1
tail -4 $OUT/gpu_tests_$TAG.log
timeout 120 python __graft_entry__.py --smoke 2>&1 | tail -1
timeout 400 python bench.py --steps 20 --warmup 5 > $OUT/bench_${TAG}_C2_T4.json 2> $OUT/bench_${TAG}_C2_T4.err
tail -c 300 $OUT/bench_${TAG}_C2_T4.err
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_${TAG}_reference.json 2> $OUT/bench_${TAG}_reference.err
run() { name=$1; shift; timeout 300 python bench.py --no-cpu-baseline --steps 20 --warmup 5 "$@" > $OUT/bench_${TAG}_$name.json 2> $OUT/bench_${TAG}_$name.err || tail -c 400 $OUT/bench_${TAG}_$name.err; }
run C2_T1 --config C2 --T 1
run C2_T2 --config C2 --T 2
run C1 --config C1
run C1_T4 --config C1 --T 4
run C3 --config C3
run C4 --config C4
run C3_strong --config C3 --scaling strong
run C4_strong --config C4 --scaling strong
for B in 1 2 4 8 16 32 64 128 256; do run C5_B$B --config C5 --batch $B; done
python - <<PY
import json, glob
for f in sorted(glob.glob("$OUT/bench_${TAG}_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "unreadable", e); continue
    r = d.get("roofline", {})
    print(f.split("bench_${TAG}_")[1], round(d["ms_per_step"], 3), "ms", round(d["value"]), "fr/s e2e", round(d["e2e"]["value"]),
          "launches", d.get("gpu_launches"), "clk", d.get("clocks", {}).get("sm_mhz"), "| top", r.get("kernel"), round(r.get("frac", 0), 3),
          "step_frac", round(r.get("step_frac", 0), 3), "rtf", (d.get("rtf") or {}).get("rtf_ref_p_rtf_cm"))
PY
# launch lists (2 steps each; the second is warm)
for cfg in C2 C3; do
  timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $OUT/launches_${TAG}_$cfg.csv \
      python tools/stage_only.py --stage all --config $cfg --reps 2 > $OUT/ncu_${TAG}_$cfg.log 2>&1
  python tools/launch_summary.py $OUT/launches_${TAG}_$cfg.csv > $OUT/launches_${TAG}_${cfg}_summary.txt; head -14 $OUT/launches_${TAG}_${cfg}_summary.txt
done
# ncu --set full: (1) one denoiser evaluation's tcgen05 kernels (skip dpen + the first evaluation), (2) the vocoder, (3) row kernels
timeout 300 ncu --set full --clock-control none --import-source on -k regex:umma_ -s 70 -c 12 -o $OUT/prof_dn_$TAG -f \
    python tools/stage_only.py --stage sampler --config C2 --T 2 --reps 1 > $OUT/prof_dn_$TAG.log 2>&1
python tools/ncu_summary.py $OUT/prof_dn_$TAG.ncu-rep $OUT/ncu_${TAG}_denoiser_summary.csv
timeout 500 ncu --set full --clock-control none -k regex:umma_ -s 64 -c 64 -o $OUT/prof_voc_$TAG -f \
    python tools/stage_only.py --stage vocoder --config C2 --reps 2 > $OUT/prof_voc_$TAG.log 2>&1
python tools/ncu_summary.py $OUT/prof_voc_$TAG.ncu-rep $OUT/ncu_${TAG}_vocoder_summary.csv
rm -f $OUT/prof_voc_$TAG.ncu-rep
timeout 300 ncu --set full --clock-control none -k "regex:length_regulate|conv_post|f32_to_f16|renoise|layernorm|round_durations|energy_embed|cwt_|embed_tokens|attention|few_rows|pack_rows" -c 60 -o $OUT/prof_rows_$TAG -f \
    python tools/stage_only.py --stage all --config C2 --T 2 --reps 1 > $OUT/prof_rows_$TAG.log 2>&1
python tools/ncu_summary.py $OUT/prof_rows_$TAG.ncu-rep $OUT/ncu_${TAG}_rowkernels_summary.csv
rm -f $OUT/prof_rows_$TAG.ncu-rep
ls -la $OUT/prof_dn_$TAG.ncu-rep; du -sh $OUT
