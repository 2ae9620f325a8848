#!/bin/bash
# A/B of the temporal-decoder row savings (tempura.DEC_FIRST_ON_PAIRS / DEC_LATTER_ONLY): parity tests, then the
# headline step with the dense schedule, each saving alone and both, inside ONE gpurun call (same box, same hour).
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_kernels_gpu.py tests/test_tempura_gpu.py tests/test_fullsize_gpu.py \
    tests/test_maskconv_gpu.py -m gpu -q > gpurun_out/r02_dec_savings_tests.log 2>&1
echo "rc=$?" >> gpurun_out/r02_dec_savings_tests.log
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-library-baseline --no-teatgt --no-e2e"
for cfg in "0 0" "0 1" "1 0" "1 1" "0 0" "1 1"; do
    set -- $cfg
    B200VSGG_DEC_FIRST_ON_PAIRS=$1 B200VSGG_DEC_LATTER_ONLY=$2 timeout 120 $B \
        > gpurun_out/r02_dec_savings_bench_$1$2.json 2>> gpurun_out/r02_dec_savings_bench.err
    python - "$1$2" <<'PY'
import json, sys
try:
    d = json.loads(open("gpurun_out/r02_dec_savings_bench_%s.json" % sys.argv[1]).read().strip().splitlines()[-1])
    r = d["roofline"]
    print("first_on_pairs,latter_only=%s  %.2f ms/step  %.1f k pairs/s  gemm %.2f ms (%.0f TF/s, %.2f TF/step)  launches %d" % (
        sys.argv[1], d["ms_per_step"], d["value"] / 1e3, r["gemm_ms_per_step"], r["achieved"], r["gemm_flops_per_step"] / 1e12, d["gpu_launches"]))
except Exception as ex:
    print(sys.argv[1], "failed:", ex)
PY
done | tee gpurun_out/r02_dec_savings_ab.txt
tail -5 gpurun_out/r02_dec_savings_tests.log
