#!/bin/bash
# Final round-2 record of the tree with the decoder row savings: the default bench line (what the driver runs) and the
# ncu launch list of one step of the same command (taken only after the plain command exited 0).
mkdir -p gpurun_out
timeout 420 python bench.py > gpurun_out/r02_bench_final_v2.json 2> gpurun_out/r02_bench_final_v2.err
echo "bench rc=$?"
CMD="python bench.py --steps 2 --profile"
timeout 120 $CMD > gpurun_out/plain_v4.log 2>&1 || exit 1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file gpurun_out/r02_launches_step_v4.csv $CMD > gpurun_out/ncu_launches_v4.log 2>&1
echo "launch list rc=$?"
python tools/summarize_launches.py gpurun_out/r02_launches_step_v4.csv gpurun_out/r02_launches_step_v4.txt | head -12
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_bench_final_v2.json").read().strip().splitlines()[-1])
print("value %.1f k pairs/s, %.2f ms/step, e2e %.1f k / %.1f k, gemm frac %.3f, model_tflops %.0f" % (
    d["value"] / 1e3, d["ms_per_step"], d["e2e"]["value"] / 1e3, d["e2e_producer_contract"]["value"] / 1e3,
    d["roofline"]["frac"], d["model_tflops"]))
print("clocks", d["clocks"])
t = d.get("teatgt") or {}
for k in ("sgcls", "predcls"):
    if k in t:
        print(k, "%.1f k pairs/s %.1f ms" % (t[k]["value"] / 1e3, t[k]["ms_per_step"]))
PY
