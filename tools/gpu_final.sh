#!/bin/bash
# round-1 final GPU validation: full GPU test suite, smoke, bench (both arms), ncu launch list of one step,
# ncu --set full captures of the conv and FDS kernels
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/ -x -q -m gpu ) > gpurun_out/final_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/final_pytest.log
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/final_smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/final_smoke.log
( timeout 600 python bench.py --steps 20 --warmup 5 ) > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err
echo "bench rc=$?" >> gpurun_out/final_bench.err
( timeout 300 python bench.py --impl reference --steps 6 --warmup 1 ) > gpurun_out/final_bench_ref.json 2> gpurun_out/final_bench_ref.err
( timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1800 --csv \
    --log-file gpurun_out/launches_r1b.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline ) > gpurun_out/final_ncu.log 2>&1
echo "ncu rc=$?" >> gpurun_out/final_ncu.log
( timeout 240 ncu --set full --clock-control none --import-source on -k regex:igemm_kernel --launch-skip 230 -c 8 -f \
    -o gpurun_out/prof_igemm_r1b python bench.py --steps 1 --warmup 3 --no-cpu-baseline ) > gpurun_out/final_ncu_igemm.log 2>&1
echo "ncu igemm rc=$?" >> gpurun_out/final_ncu_igemm.log
( timeout 240 ncu --set full --clock-control none --import-source on -k regex:fds_accumulate --launch-skip 4 -c 3 -f \
    -o gpurun_out/prof_fds_r1b python bench.py --steps 1 --warmup 3 --no-cpu-baseline ) > gpurun_out/final_ncu_fds.log 2>&1
echo "ncu fds rc=$?" >> gpurun_out/final_ncu_fds.log
tail -4 gpurun_out/final_pytest.log; tail -2 gpurun_out/final_smoke.log; cut -c1-400 gpurun_out/final_bench.json; tail -2 gpurun_out/final_bench.err; cut -c1-300 gpurun_out/final_bench_ref.json; tail -2 gpurun_out/final_ncu.log; tail -2 gpurun_out/final_ncu_igemm.log; tail -2 gpurun_out/final_ncu_fds.log
exit 0
