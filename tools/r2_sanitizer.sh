#!/bin/bash
# compute-sanitizer passes: racecheck on the conv kernels' mbarrier / TMEM protocol (every GEMM
# form incl. CTA pairs, small shapes), memcheck on the FDS / loss / LDS kernels and on one tiny ResNet training step.
# Summaries go to profiles/ (copied by hand from gpurun_out/).
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
CS=/usr/local/cuda/bin/compute-sanitizer
( timeout 900 $CS --tool racecheck --racecheck-report analysis --print-limit 20 python -m pytest tests/test_gpu_conv.py -q -x -k forms ) > gpurun_out/r2_sanitizer_racecheck_conv.log 2>&1
echo "racecheck rc=$?" >> gpurun_out/r2_sanitizer_racecheck_conv.log
( timeout 900 $CS --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_fds_loss_lds.py -q -x -k "not config4 and not config5 and not full_size" ) > gpurun_out/r2_sanitizer_memcheck_fds.log 2>&1
echo "memcheck fds rc=$?" >> gpurun_out/r2_sanitizer_memcheck_fds.log
( timeout 900 $CS --tool memcheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/r2_sanitizer_memcheck_smoke.log 2>&1
echo "memcheck smoke rc=$?" >> gpurun_out/r2_sanitizer_memcheck_smoke.log
for f in gpurun_out/r2_sanitizer_*.log; do echo "== $f"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|rc=|smoke ok" $f | tail -6; done
exit 0
