#!/bin/bash
# round-2 GPU call 9: ResNet / train-loop tests after the max-pool / BN-apply changes, bench, launch list
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
( time timeout 900 python -m pytest tests/test_gpu_resnet.py tests/test_gpu_train_loop.py tests/test_gpu_train_script.py -q 2>&1 ) > gpurun_out/r2c9_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2c9_pytest.log
( timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline ) > gpurun_out/r2c9_bench.json 2> gpurun_out/r2c9_bench.err
( timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1300 --csv \
    --log-file gpurun_out/launches_r2c.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline ) > gpurun_out/r2c9_ncu.log 2>&1
grep -E "passed|failed|FAILED" gpurun_out/r2c9_pytest.log | tail -8; cut -c1-300 gpurun_out/r2c9_bench.json; tail -3 gpurun_out/r2c9_bench.err
exit 0
