#!/bin/bash
# round-2 GPU call 4: whole GPU suite (no -x), diagnosis of the teacher-forced backward tolerance, ncu launch list of a step
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
( time timeout 1500 python -m pytest tests/ -q -m gpu 2>&1 ) > gpurun_out/r2c4_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2c4_pytest.log
T="tests/test_gpu_resnet.py::test_backward_vs_oracle_teacher_forced"
( timeout 300 python -m pytest "$T" -q -s 2>&1 | grep -E "worst|passed|failed" ) > gpurun_out/r2c4_diag_default.log 2>&1
( DIRB200_TEST_TF32=1 timeout 300 python -m pytest "$T" -q -s 2>&1 | grep -E "worst|passed|failed" ) > gpurun_out/r2c4_diag_tf32.log 2>&1
( DIRB200_IM2COL=0 DIRB200_CTA2=0 DIRB200_FUSED_STATS=0 timeout 300 python -m pytest "$T" -q -s 2>&1 | grep -E "worst|passed|failed" ) > gpurun_out/r2c4_diag_r1paths.log 2>&1
( timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1400 --csv \
    --log-file gpurun_out/launches_r2a.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline ) > gpurun_out/r2c4_ncu.log 2>&1
echo "ncu rc=$?" >> gpurun_out/r2c4_ncu.log
( timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline ) > gpurun_out/r2c4_bench.json 2> gpurun_out/r2c4_bench.err
tail -12 gpurun_out/r2c4_pytest.log; for f in gpurun_out/r2c4_diag_*.log; do echo "== $f"; cat $f; done; tail -2 gpurun_out/r2c4_ncu.log; cut -c1-400 gpurun_out/r2c4_bench.json
exit 0
