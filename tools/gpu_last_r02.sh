#!/bin/bash
# Last GPU call of round 2: the tests that cover everything changed since the last full green run (SGCls sequence plan,
# worker-thread graph build, decoder-savings tolerance), smoke, then a quick SGCls step timing with what is left.
mkdir -p gpurun_out
timeout 75 python -m pytest tests/test_sgcls_gpu.py tests/test_teatgt_gpu.py \
    "tests/test_tempura_gpu.py::test_decoder_row_savings_equal_dense_schedule" -m gpu -q -x > gpurun_out/r02_last_tests.log 2>&1
echo "tests rc=$?"; tail -2 gpurun_out/r02_last_tests.log
timeout 25 python tools/bench_teatgt.py --mode sgcls --steps 4 --warmup 2 > gpurun_out/r02_teat_sgcls_after.json 2> gpurun_out/r02_teat_sgcls_after.err
echo "sgcls rc=$?"; tail -1 gpurun_out/r02_teat_sgcls_after.json | cut -c1-330
