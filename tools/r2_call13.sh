#!/bin/bash
# round-2 GPU call 13: whole GPU suite after removing the unused kernel variants + new weight prep, bench, conv layer timings
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
( time timeout 1500 python -m pytest tests/ -q -m gpu 2>&1 ) > gpurun_out/r2c13_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2c13_pytest.log
( timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline ) > gpurun_out/r2c13_bench.json 2> gpurun_out/r2c13_bench.err
( DIRB200_TAG=r2_final timeout 300 python tests/cta2_check.py time ) > gpurun_out/r2c13_time.log 2>&1
grep -E "passed|failed|FAILED" gpurun_out/r2c13_pytest.log | tail -8; cut -c1-300 gpurun_out/r2c13_bench.json; tail -2 gpurun_out/r2c13_time.log
exit 0
