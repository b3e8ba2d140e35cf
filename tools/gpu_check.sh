#!/bin/bash
# quick GPU iteration loop used during round 1: conv + ResNet parity, per-layer conv timing, short bench
mkdir -p gpurun_out
( time timeout 400 python -m pytest tests/test_gpu_conv.py tests/test_gpu_resnet.py -q ) > gpurun_out/check_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/check_tests.log
DIRB200_CTA2=${DIRB200_CTA2:-0} timeout 120 python tests/cta2_check.py time > gpurun_out/check_layers.log 2>&1
( timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline ) > gpurun_out/check_bench.json 2> gpurun_out/check_bench.err
tail -4 gpurun_out/check_tests.log; tail -14 gpurun_out/check_layers.log; cut -c1-300 gpurun_out/check_bench.json
exit 0
