#!/bin/bash
# GPU check: conv + ResNet parity with the TMA-fed 1x1 GEMMs / narrow-staging epilogue, per-layer timing, short bench
mkdir -p gpurun_out
( time timeout 400 python -m pytest tests/test_gpu_conv.py tests/test_gpu_resnet.py -q ) > gpurun_out/c4_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/c4_tests.log
DIRB200_CTA2=0 timeout 120 python tests/cta2_check.py time > gpurun_out/c4_time_atma.log 2>&1
( timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline ) > gpurun_out/c4_bench.json 2> gpurun_out/c4_bench.err
tail -4 gpurun_out/c4_tests.log; cat gpurun_out/c4_time_atma.log | tail -14; cut -c1-300 gpurun_out/c4_bench.json; tail -3 gpurun_out/c4_bench.err
exit 0
