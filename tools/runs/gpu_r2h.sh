#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/step_timeline.py tf32 bf16 > gpurun_out/r2h_timeline_tf32_bf16.txt 2> gpurun_out/r2h_timeline.err; echo "rc=$?"
timeout 300 python tools/step_timeline.py bf16 bf16 > gpurun_out/r2h_timeline_bf16_bf16.txt 2>> gpurun_out/r2h_timeline.err; echo "rc=$?"
tail -n 3 gpurun_out/r2h_timeline.err; head -3 gpurun_out/r2h_timeline_tf32_bf16.txt
