#!/bin/bash
# compute-sanitizer over the smallest programs that exercise every hand-rolled synchronisation protocol of the library
# (mbarrier rings + tcgen05 commit in gemm_tc.cu / conv_tc.cu / attention_tc128.cu, split-K red.global.add, the side
# stream fork/join): __graft_entry__.smoke() (B = 4 train step in fp32 / tf32 / bf16 + LBS) and a B = 8 train step
# with both seam dtypes, the coarse forward and the LBS backward.  ONE tool per gpurun call (B200_PROFILING.md):
#   gpurun -- 'bash tools/sanitize.sh memcheck'     then, in another call,     gpurun -- 'bash tools/sanitize.sh racecheck'
# Summaries land in gpurun_out/sanitize_<tool>.txt; copy them to profiles/.
TOOL=${1:-memcheck}
mkdir -p gpurun_out
OUT=gpurun_out/sanitize_${TOOL}.txt
: > $OUT
run() {
  echo "== compute-sanitizer --tool $TOOL $*" >> $OUT
  timeout 1500 compute-sanitizer --tool $TOOL --print-limit 20 "$@" > gpurun_out/sanitize_${TOOL}_raw.log 2>&1
  echo "exit code $?" >> $OUT
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|Invalid|hazard|error" gpurun_out/sanitize_${TOOL}_raw.log | sort | uniq -c | head -20 >> $OUT
  grep -E "^\[smoke\]|sanitize-step" gpurun_out/sanitize_${TOOL}_raw.log >> $OUT
}
run python -c "import __graft_entry__ as g; g.smoke()"
run python tools/sanitize_step.py
cat $OUT
