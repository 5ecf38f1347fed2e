#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_lbs.py -x -q > gpurun_out/r2i_pytest_lbs.log 2>&1; echo "pytest lbs rc=$?" > gpurun_out/r2i_rc.log
timeout 300 python tools/step_timeline.py tf32 bf16 > gpurun_out/r2i_timeline_tf32_bf16.txt 2> gpurun_out/r2i_timeline.err; echo "timeline rc=$?" >> gpurun_out/r2i_rc.log
timeout 300 python tools/step_timeline.py bf16 bf16 > gpurun_out/r2i_timeline_bf16_bf16.txt 2>> gpurun_out/r2i_timeline.err
timeout 1200 python -m pytest tests -q -m gpu > gpurun_out/r2i_pytest_all.log 2>&1; echo "pytest all rc=$?" >> gpurun_out/r2i_rc.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err; echo "bench rc=$?" >> gpurun_out/r2i_rc.log
cat gpurun_out/r2i_rc.log; tail -n 6 gpurun_out/r2i_pytest_lbs.log; tail -n 6 gpurun_out/r2i_pytest_all.log; tail -n 4 gpurun_out/r2i_bench.err; head -3 gpurun_out/r2i_timeline_tf32_bf16.txt
