#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_head.py tests/test_gpu_optim.py -x -q > gpurun_out/r2e_pytest.log 2>&1; echo "pytest rc=$?" > gpurun_out/r2e_rc.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err; echo "bench rc=$?" >> gpurun_out/r2e_rc.log
cat gpurun_out/r2e_rc.log; tail -n 5 gpurun_out/r2e_pytest.log; tail -n 5 gpurun_out/r2e_bench.err
