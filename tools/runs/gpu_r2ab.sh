#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/exchange_probe.py > gpurun_out/r2ab_probe_default.log 2>&1
SCAT_EXP_ATTN_L1=1 timeout 300 python tools/exchange_probe.py > gpurun_out/r2ab_probe_attn_l1.log 2>&1
timeout 300 python tools/exchange_probe.py > gpurun_out/r2ab_probe_default2.log 2>&1
SCAT_EXP_ATTN_L1=1 timeout 300 python tools/exchange_probe.py > gpurun_out/r2ab_probe_attn_l1_2.log 2>&1
grep -H EXCHANGE_PROBE gpurun_out/r2ab_probe_*.log
