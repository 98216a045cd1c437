#!/bin/bash
# ncu evidence for one shortened step (large-v3, B=64): launch list + full captures of the top kernels of every stage.
# Kernel nodes of the decode CUDA graph are profiled individually (ncu's default graph-profiling mode).
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
    python scripts/profile_step.py --batch 64 --tokens 4 --vad --align --beam 2 --scenes > gpurun_out/prof_launch.log 2>&1
ncu --set full --clock-control none -k regex:"gemm_tc|attn_encoder|logmel_kernel|layernorm" -c 14 \
    -o gpurun_out/prof_encoder -f python scripts/profile_step.py --batch 64 --tokens 2 > gpurun_out/prof_enc.log 2>&1
ncu --set full --clock-control none -k regex:"attn_dec|gemm_step|sample_kernel" -s 44 -c 14 \
    -o gpurun_out/prof_decode -f python scripts/profile_step.py --batch 64 --tokens 3 > gpurun_out/prof_dec.log 2>&1
ncu --set full --clock-control none -k regex:"vad_|align_|scene_|beam_select" -c 10 \
    -o gpurun_out/prof_vad_align -f python scripts/profile_step.py --batch 64 --tokens 12 --vad --align --beam 2 --scenes > gpurun_out/prof_va.log 2>&1
# gpurun copies back at most 64 MiB: keep the raw metric pages (CSV), drop the reports
for r in gpurun_out/prof_*.ncu-rep; do
  ncu -i "$r" --page raw --csv > "${r%.ncu-rep}.raw.csv" 2>/dev/null
  rm -f "$r"
done
ls -la gpurun_out/*.raw.csv gpurun_out/launches.csv
