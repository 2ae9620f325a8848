#!/bin/bash
# Round profile pass: launch list of one step + ncu --set full of the dominant kernels (each only after the same
# command exited 0 without ncu).
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --profile"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
$CMD > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"gemm2_bf16_kernel|gemm_bf16_kernel|attn_mma_bwd|attn_mma_fwd|layernorm_bwd" -s 900 -c 60 -o gpurun_out/prof_step $CMD > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"
ncu -i gpurun_out/prof_step.ncu-rep --page raw --csv > gpurun_out/prof_step_raw.csv 2>/dev/null
echo "raw csv rc=$? $(wc -c < gpurun_out/prof_step_raw.csv) bytes"
