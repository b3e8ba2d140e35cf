#!/bin/bash
# round-2 GPU call 3: full GPU suite after the conv / BN rework, conv layer timing, short bench (fused stats on / off)
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
( time timeout 1500 python -m pytest tests/ -q -m gpu -x --deselect tests/test_gpu_conv_variants.py 2>&1 ) > gpurun_out/r2c3_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2c3_pytest.log
( DIRB200_TAG=r2c3 timeout 300 python tests/cta2_check.py time ) > gpurun_out/r2c3_time.log 2>&1
( timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline ) > gpurun_out/r2c3_bench.json 2> gpurun_out/r2c3_bench.err
( DIRB200_FUSED_STATS=0 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline ) > gpurun_out/r2c3_bench_nofuse.json 2> gpurun_out/r2c3_bench_nofuse.err
tail -15 gpurun_out/r2c3_pytest.log; tail -2 gpurun_out/r2c3_time.log; cut -c1-600 gpurun_out/r2c3_bench.json; tail -3 gpurun_out/r2c3_bench.err; cut -c1-300 gpurun_out/r2c3_bench_nofuse.json
exit 0
