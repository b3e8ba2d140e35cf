#!/bin/bash
# round-2 GPU call 41 (2 GPUs): raw NCCL time of the 94 MB gradient all-reduce under a few NCCL settings
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
run() { ( env "$@" timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 tools/nccl_allreduce_bench.py ) >> gpurun_out/r2c41_nccl.log 2>> gpurun_out/r2c41_nccl.err; }
rm -f gpurun_out/r2c41_nccl.log gpurun_out/r2c41_nccl.err
run X=1
run NCCL_MIN_NCHANNELS=32
run NCCL_PROTO=Simple
run NCCL_ALGO=Ring NCCL_PROTO=Simple NCCL_MIN_NCHANNELS=32 NCCL_NTHREADS=512
run NCCL_BUFFSIZE=16777216 NCCL_MIN_NCHANNELS=32
run NCCL_P2P_USE_CUDA_MEMCPY=1
cat gpurun_out/r2c41_nccl.log; grep -i "error\|warn" gpurun_out/r2c41_nccl.err | head -5
exit 0
