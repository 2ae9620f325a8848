#!/bin/bash
# launch list of one step (ncu, serialised) -> gpurun_out/launches.csv
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --profile"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches rc=$?"
