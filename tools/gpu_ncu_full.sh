#!/bin/bash
# ncu --set full of the dominant kernels of one bench step; reports are reduced to raw CSV on the box
# (the .ncu-rep files are too large to travel back).
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --profile"
$CMD > gpurun_out/plain2.log 2>&1 || exit 1
run() {  # name regex skip count
  ncu --set full --clock-control none -k regex:$2 -s $3 -c $4 -o /tmp/prof_$1 $CMD > gpurun_out/ncu_$1.log 2>&1
  ncu -i /tmp/prof_$1.ncu-rep --page raw --csv > gpurun_out/ncu_$1_raw.csv 2>/dev/null
  echo "$1 rc=$? $(wc -c < gpurun_out/ncu_$1_raw.csv) bytes"
}
run gemm gemm_bf16_kernel 100 24
run attn attn_mma 8 8
run rows "layernorm_bwd|seg_colstats|nchw_to_nhwc_bf16|adamw_clip" 10 8
