#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_head.py -x -q > gpurun_out/r2d_pytest.log 2>&1; echo "pytest rc=$?" > gpurun_out/r2d_rc.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err; echo "bench rc=$?" >> gpurun_out/r2d_rc.log
timeout 300 python bench.py --steps 20 --warmup 5 --quick --seam fp32 > gpurun_out/r2d_bench_fp32seam.json 2>> gpurun_out/r2d_bench.err; echo "bench fp32 seam rc=$?" >> gpurun_out/r2d_rc.log
python bench.py --steps 2 --warmup 3 --quick > /dev/null 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2d_launches.csv python bench.py --steps 2 --warmup 3 --quick > gpurun_out/r2d_ncu.log 2>&1
python tools/summarize_launches.py gpurun_out/r2d_launches.csv > gpurun_out/r2d_launches.txt 2>&1
cat gpurun_out/r2d_rc.log; tail -n 5 gpurun_out/r2d_pytest.log; tail -n 5 gpurun_out/r2d_bench.err; head -45 gpurun_out/r2d_launches.txt
