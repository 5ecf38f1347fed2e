#!/bin/bash
mkdir -p gpurun_out
for k in 1 2; do
SCAT_SAVE_DGELU=0 timeout 300 python tools/exchange_probe.py > gpurun_out/r2af_probe_off$k.log 2>&1
timeout 300 python tools/exchange_probe.py > gpurun_out/r2af_probe_on$k.log 2>&1
done
grep -H EXCHANGE_PROBE gpurun_out/r2af_probe_*.log
