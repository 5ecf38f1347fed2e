#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc=$?" > gpurun_out/r2f_rc.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2f_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2f_rc.log
cat gpurun_out/r2f_rc.log; tail -n 25 gpurun_out/r2f_pytest.log; tail -n 6 gpurun_out/r2f_smoke.log
