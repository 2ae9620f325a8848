#!/bin/bash
# ncu --set full of the tcgen05 TokenGT attention kernels (forward, dQ, dK/dV) at the training-batch shape (448 clips of
# 250..589 tokens, 32 heads x 24, dropout 0.1) — run only after the same command exited 0 without ncu.
tag=${1:-r2}; shape=${2:-short}
mkdir -p gpurun_out
python tools/attn_tc_profile.py 0.1 $shape > gpurun_out/${tag}_attn_tc_bench.txt 2>&1 || exit 1
cat gpurun_out/${tag}_attn_tc_bench.txt
ncu --set full --clock-control none --import-source on -k regex:attn_tc -s 3 -c 3 -f -o gpurun_out/${tag}_attn_tc \
    python tools/attn_tc_profile.py 0.1 $shape > gpurun_out/${tag}_ncu_attn_tc.log 2>&1
ncu -i gpurun_out/${tag}_attn_tc.ncu-rep --page raw --csv > gpurun_out/${tag}_attn_tc_raw.csv 2>/dev/null
ncu -i gpurun_out/${tag}_attn_tc.ncu-rep --page source --csv > gpurun_out/${tag}_attn_tc_source.csv 2>/dev/null
ls -la gpurun_out/${tag}_attn_tc*
