#!/bin/bash
# GPU check of the im2col-TMA A operand (DIRB200_IM2COL=1): conv parity, per-layer timing, ResNet parity, short bench
mkdir -p gpurun_out
export DIRB200_IM2COL=1
( time timeout 200 python tests/cta2_check.py parity ) > gpurun_out/i2c_parity.log 2>&1
rc=$?
echo "im2col parity rc=$rc" >> gpurun_out/i2c_parity.log
if [ $rc -eq 0 ]; then
  timeout 120 python tests/cta2_check.py time > gpurun_out/i2c_time.log 2>&1
  ( time timeout 400 python -m pytest tests/test_gpu_conv.py tests/test_gpu_resnet.py -q ) > gpurun_out/i2c_tests.log 2>&1
  echo "tests rc=$?" >> gpurun_out/i2c_tests.log
  ( timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline ) > gpurun_out/i2c_bench.json 2> gpurun_out/i2c_bench.err
fi
grep -E "PASS|FAIL|parity" gpurun_out/i2c_parity.log | tail -14; tail -13 gpurun_out/i2c_time.log 2>/dev/null; tail -4 gpurun_out/i2c_tests.log 2>/dev/null; cut -c1-300 gpurun_out/i2c_bench.json 2>/dev/null
exit 0
