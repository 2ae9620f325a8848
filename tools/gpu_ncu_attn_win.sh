#!/bin/bash
# ncu --set full of the window-attention kernels at the headline shape (31.8 k window tokens, 8 heads x 242, dropout 0.1):
# run only after the same command exited 0 without ncu.  Output: gpurun_out/<tag>_attn_win.ncu-rep + raw/source CSV pages.
tag=${1:-r2}
mkdir -p gpurun_out
python tools/attn_bench.py > gpurun_out/${tag}_attn_bench.txt 2>&1 || exit 1
cat gpurun_out/${tag}_attn_bench.txt
ncu --set full --clock-control none --import-source on -k regex:attn_win -s 68 -c 2 -f -o gpurun_out/${tag}_attn_win \
    python tools/attn_bench.py > gpurun_out/${tag}_ncu_attn_win.log 2>&1
ncu -i gpurun_out/${tag}_attn_win.ncu-rep --page raw --csv > gpurun_out/${tag}_attn_win_raw.csv 2>/dev/null
ncu -i gpurun_out/${tag}_attn_win.ncu-rep --page source --csv > gpurun_out/${tag}_attn_win_source.csv 2>/dev/null
ls -la gpurun_out/${tag}_attn_win*
