#!/bin/bash
# GPU check: conv + ResNet parity after the wgrad split search / BN tile rule / fused BN partial sums, short bench
mkdir -p gpurun_out
( time timeout 400 python -m pytest tests/test_gpu_conv.py tests/test_gpu_resnet.py -q ) > gpurun_out/c5_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/c5_tests.log
( timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline ) > gpurun_out/c5_bench.json 2> gpurun_out/c5_bench.err
tail -4 gpurun_out/c5_tests.log; cut -c1-300 gpurun_out/c5_bench.json; tail -3 gpurun_out/c5_bench.err
exit 0
