#!/bin/bash
# round-2 GPU call 7 (2 GPUs): multi-GPU correctness check + 2-rank bench (overlapped gradient all-reduce)
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
( timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tests/mgpu_check.py ) > gpurun_out/r2c7_mgpu.log 2>&1
echo "mgpu rc=$?" >> gpurun_out/r2c7_mgpu.log
( NCCL_DEBUG=INFO timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --steps 20 --warmup 5 ) > gpurun_out/r2c7_bench2.json 2> gpurun_out/r2c7_bench2.err
echo "bench2 rc=$?" >> gpurun_out/r2c7_bench2.err
( timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline ) > gpurun_out/r2c7_bench1.json 2> gpurun_out/r2c7_bench1.err
tail -5 gpurun_out/r2c7_mgpu.log; wc -l gpurun_out/r2c7_bench2.json; cut -c1-250 gpurun_out/r2c7_bench2.json; grep -c "NCCL INFO" gpurun_out/r2c7_bench2.err; grep -E "nranks|NVLS|bench2 rc" gpurun_out/r2c7_bench2.err | head -5; cut -c1-250 gpurun_out/r2c7_bench1.json
exit 0
