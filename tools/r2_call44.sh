#!/bin/bash
# round-2 GPU call 44: train.py with --device_transform (uint8 batches, GPU crop / flip / normalise), input-pipeline tests
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
( timeout 600 python -m pytest tests/test_gpu_train_script.py tests/test_gpu_input_pipeline.py tests/test_gpu_train_loop.py -q 2>&1 | tail -8 ) > gpurun_out/r2c44_pytest.log 2>&1
cat gpurun_out/r2c44_pytest.log
exit 0
