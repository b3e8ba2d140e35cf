#!/bin/bash
# round-2 GPU call 12 (2 GPUs): ResNet tests after the prep-kernel change, multi-GPU check, 2-rank bench with NCCL CTA limits
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
( timeout 600 python -m pytest tests/test_gpu_resnet.py tests/test_gpu_conv.py -q 2>&1 | tail -3 ) > gpurun_out/r2c12_pytest.log 2>&1
( timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tests/mgpu_check.py ) > gpurun_out/r2c12_mgpu.log 2>&1
echo "mgpu rc=$?" >> gpurun_out/r2c12_mgpu.log
( timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline ) > gpurun_out/r2c12_bench1.json 2> gpurun_out/r2c12_bench1.err
for ctas in default 4 8 16; do
  if [ $ctas = default ]; then E="X=1"; else E="NCCL_MAX_CTAS=$ctas"; fi
  ( env $E timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --steps 20 --warmup 5 ) > gpurun_out/r2c12_bench2_$ctas.json 2> gpurun_out/r2c12_bench2_$ctas.err
done
cat gpurun_out/r2c12_pytest.log; tail -3 gpurun_out/r2c12_mgpu.log; for f in gpurun_out/r2c12_bench1.json gpurun_out/r2c12_bench2_*.json; do echo "== $f"; cut -c1-200 $f; done
exit 0
