#!/bin/bash
# round-2 GPU call 5: GPU suite + bench after the epilogue-stats / BN-grid / staged-backward / FDS-small changes
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
( time timeout 1500 python -m pytest tests/ -q -m gpu 2>&1 ) > gpurun_out/r2c5_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2c5_pytest.log
( timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline ) > gpurun_out/r2c5_bench.json 2> gpurun_out/r2c5_bench.err
( DIRB200_FUSED_STATS=0 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline ) > gpurun_out/r2c5_bench_nofuse.json 2> gpurun_out/r2c5_bench_nofuse.err
( timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1400 --csv \
    --log-file gpurun_out/launches_r2b.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline ) > gpurun_out/r2c5_ncu.log 2>&1
grep -E "passed|failed|FAILED" gpurun_out/r2c5_pytest.log | tail -12; cut -c1-300 gpurun_out/r2c5_bench.json; tail -3 gpurun_out/r2c5_bench.err; cut -c1-200 gpurun_out/r2c5_bench_nofuse.json
exit 0
