#!/bin/bash
# round-2 GPU call 2: conv kernels after removing the run-time divisions from the producer / TMA k-loops
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
( timeout 400 python -m pytest tests/test_gpu_conv.py -q -x ) > gpurun_out/r2c2_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2c2_tests.log
( DIRB200_TAG=_nodiv timeout 300 python tests/cta2_check.py time ) > gpurun_out/r2c2_time_default.log 2>&1
for m in "DIRB200_CTA2=2 DIRB200_IM2COL=1" "DIRB200_IM2COL=1" "DIRB200_CTA2=3" "DIRB200_BSTAT=1"; do
  tag=$(echo $m | tr -d ' =A-Z_')
  ( env $m timeout 300 python tests/cta2_check.py parity ) > gpurun_out/r2c2_parity_$tag.log 2>&1
  echo "parity rc=$?" >> gpurun_out/r2c2_parity_$tag.log
  ( env $m DIRB200_TAG=_nodiv_$tag timeout 300 python tests/cta2_check.py time ) > gpurun_out/r2c2_time_$tag.log 2>&1
done
tail -3 gpurun_out/r2c2_tests.log
for f in gpurun_out/r2c2_parity_*.log; do echo "== $f"; grep -E "FAIL|parity" $f | tail -3; done
for f in gpurun_out/r2c2_time_*.log; do echo "== $f"; tail -1 $f; done
exit 0
