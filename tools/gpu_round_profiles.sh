#!/bin/bash
# Round profile pass: launch list of one step + ncu --set full of the dominant kernels (each only after the same
# command exited 0 without ncu).  --launch-skip counts MATCHING launches; a step has ~90 GEMM launches.
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --profile"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
run() {  # name regex skip count
  ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c $4 -o /tmp/prof_$1 $CMD > gpurun_out/ncu_$1.log 2>&1
  ncu -i /tmp/prof_$1.ncu-rep --page raw --csv > gpurun_out/ncu_$1_raw.csv 2>/dev/null
  echo "$1 rc=$? $(wc -c < gpurun_out/ncu_$1_raw.csv) bytes"
}
$CMD > gpurun_out/plain2.log 2>&1 || exit 1
run gemm "gemm2_bf16_kernel|gemm_bf16_kernel" 190 30
run attn "attn_mma" 18 8
run rows "layernorm_bwd_dx|layernorm_fwd|seg_colstats|gated_residual" 100 12
