#!/bin/bash
# round-2 GPU call 33: BN / pooling entry points, refinement module R, config-4 figures after the FDS CTA change
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
( timeout 900 python -m pytest tests/test_gpu_dense_ops.py tests/test_gpu_fds_loss_lds.py -q 2>&1 | tail -15 ) > gpurun_out/r2c33_pytest.log 2>&1
( timeout 300 python tools/nyud2_bench.py ) > gpurun_out/r2c33_nyud2.json 2> gpurun_out/r2c33_nyud2.err
cat gpurun_out/r2c33_pytest.log; cat gpurun_out/r2c33_nyud2.json; tail -3 gpurun_out/r2c33_nyud2.err
exit 0
