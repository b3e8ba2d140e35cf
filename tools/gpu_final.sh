#!/bin/bash
# round-1 final GPU validation (trimmed to the GPU minutes left): full GPU test suite, bench, ncu launch list of one
# step, smoke.  (`bench.py --impl reference` is CPU-only and was run separately.)
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/ -x -q -m gpu ) > gpurun_out/final_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/final_pytest.log
( timeout 300 python bench.py --steps 20 --warmup 5 ) > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err
echo "bench rc=$?" >> gpurun_out/final_bench.err
( timeout 240 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1250 --csv \
    --log-file gpurun_out/launches_r1b.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline ) > gpurun_out/final_ncu.log 2>&1
echo "ncu rc=$?" >> gpurun_out/final_ncu.log
( timeout 120 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/final_smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/final_smoke.log
tail -4 gpurun_out/final_pytest.log; cut -c1-400 gpurun_out/final_bench.json; tail -2 gpurun_out/final_bench.err; tail -2 gpurun_out/final_ncu.log; tail -2 gpurun_out/final_smoke.log
exit 0
