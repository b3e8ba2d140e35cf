#!/bin/bash
# round-2 final validation (1 GPU): whole GPU suite, smoke, bench with the CPU arm, reference arm, ncu launch list
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
( time timeout 1500 python -m pytest tests/ -q -m gpu 2>&1 | tail -6 ) > gpurun_out/r2final_pytest.log 2>&1
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/r2final_smoke.log 2>&1
( timeout 600 python bench.py --steps 30 --warmup 5 ) > gpurun_out/r2final_bench.json 2> gpurun_out/r2final_bench.err
( timeout 600 python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/r2final_bench_ref.json 2> gpurun_out/r2final_bench_ref.err
( timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1400 --csv \
    --log-file gpurun_out/launches_r2final.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline ) > gpurun_out/r2final_ncu.log 2>&1
( DIRB200_TAG=r2_final timeout 300 python tests/cta2_check.py time ) > gpurun_out/r2final_time.log 2>&1
cat gpurun_out/r2final_pytest.log; tail -1 gpurun_out/r2final_smoke.log; cut -c1-250 gpurun_out/r2final_bench.json; cut -c1-200 gpurun_out/r2final_bench_ref.json; tail -1 gpurun_out/r2final_time.log
exit 0
