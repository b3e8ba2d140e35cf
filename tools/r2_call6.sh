#!/bin/bash
# round-2 GPU call 6: GPU suite + bench with programmatic dependent launch on / off, FDS small-sort path
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
( time timeout 1500 python -m pytest tests/ -q -m gpu -x 2>&1 ) > gpurun_out/r2c6_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2c6_pytest.log
( timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline ) > gpurun_out/r2c6_bench.json 2> gpurun_out/r2c6_bench.err
( DIRB200_PDL=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline ) > gpurun_out/r2c6_bench_nopdl.json 2> gpurun_out/r2c6_bench_nopdl.err
grep -E "passed|failed|FAILED" gpurun_out/r2c6_pytest.log | tail -12; cut -c1-300 gpurun_out/r2c6_bench.json; tail -3 gpurun_out/r2c6_bench.err; cut -c1-200 gpurun_out/r2c6_bench_nopdl.json
exit 0
