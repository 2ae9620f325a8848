#!/bin/bash
# ncu --set full (with source correlation) of single GEMM shapes; reduced to CSV on the box.
mkdir -p gpurun_out
run() {  # name, gemm_one args...
  name=$1; shift
  python tools/gemm_one.py "$@" > gpurun_out/gemm_one_$name.log 2>&1 || { echo "$name plain run failed"; return; }
  cat gpurun_out/gemm_one_$name.log
  ncu --set full --import-source on --clock-control none -k regex:gemm -s 3 -c 1 -o /tmp/p_$name python tools/gemm_one.py "$@" > gpurun_out/ncu_$name.log 2>&1
  ncu -i /tmp/p_$name.ncu-rep --page raw --csv > gpurun_out/ncu_${name}_raw.csv 2>/dev/null
  ncu -i /tmp/p_$name.ncu-rep --page source --csv > gpurun_out/ncu_${name}_src.csv 2>/dev/null
  echo "$name raw=$(wc -c < gpurun_out/ncu_${name}_raw.csv) src=$(wc -c < gpurun_out/ncu_${name}_src.csv)"
}
run smallk 804531 1152 256 0 1
run resf32 31799 1936 1936 0 0 f32res
run ffn 18467 7744 1936 0 0
