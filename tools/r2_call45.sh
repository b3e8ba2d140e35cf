#!/bin/bash
# round-2 GPU call 45: LDS / dataset tests after the reweight == 'none' early return
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
( timeout 300 python -m pytest tests/test_gpu_fds_loss_lds.py tests/test_gpu_train_script.py -q -k "lds or LDS or weights or script" 2>&1 | tail -5 ) > gpurun_out/r2c45_pytest.log 2>&1
cat gpurun_out/r2c45_pytest.log
exit 0
