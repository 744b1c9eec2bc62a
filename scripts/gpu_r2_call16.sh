#!/bin/bash
# full GPU suite + default bench + smoke on the current build
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/r2c16_tests.log 2>&1; echo "tests rc=$?" > gpurun_out/r2c16_summary.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2c16_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2c16_summary.txt
timeout 900 python bench.py > gpurun_out/r2c16_bench.json 2> gpurun_out/r2c16_bench.err; echo "bench rc=$?" >> gpurun_out/r2c16_summary.txt
timeout 600 python bench.py --impl reference > gpurun_out/r2c16_bench_ref.json 2> gpurun_out/r2c16_bench_ref.err; echo "ref rc=$?" >> gpurun_out/r2c16_summary.txt
tail -8 gpurun_out/r2c16_tests.log; cat gpurun_out/r2c16_smoke.log | tail -2; cat gpurun_out/r2c16_summary.txt
