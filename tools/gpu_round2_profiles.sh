#!/bin/bash
# Round-2 evidence pass (one B200): every ncu command runs only after the same command exited 0 without ncu.
#   1. launch list of one TEMPURA training step        -> gpurun_out/r02_launches_step.csv
#   2. ncu --set full: window attention (fwd + bwd), tcgen05 TokenGT attention (fwd, dQ, dK/dV), GEMMs + row kernels in-step
#   3. launch list of one TEAT-GT PredCLS training step -> gpurun_out/r02_teat_launches.csv
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --profile"
$CMD > gpurun_out/plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file gpurun_out/r02_launches_step.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
bash tools/gpu_ncu_attn_win.sh r02
bash tools/gpu_ncu_attn_tc.sh r02 short
run() {  # name regex skip count
  ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c $4 -f -o gpurun_out/r02_$1 $CMD > gpurun_out/ncu_$1.log 2>&1
  ncu -i gpurun_out/r02_$1.ncu-rep --page raw --csv > gpurun_out/r02_$1_raw.csv 2>/dev/null
  echo "$1 rc=$? $(wc -c < gpurun_out/r02_$1_raw.csv) bytes"
}
run gemm "gemm2_bf16_kernel|gemm_bf16_kernel" 190 24
run rows "layernorm_bwd_dx|layernorm_fwd|seg_colstats|colsum_vec|gated_residual|graph_small|cast_dropout" 120 14
TCMD="python tools/bench_teatgt.py --steps 1 --warmup 1"
$TCMD > gpurun_out/plain_teat.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r02_teat_launches.csv $TCMD > gpurun_out/ncu_teat.log 2>&1
echo "teat launch list rc=$?"
rm -f gpurun_out/r02_gemm.ncu-rep gpurun_out/r02_rows.ncu-rep     # the raw CSV pages are kept; the reports are large
