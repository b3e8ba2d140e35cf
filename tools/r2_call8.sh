#!/bin/bash
# round-2 GPU call 8: compute-sanitizer passes + ncu --set full of the layer-1 BN backward kernels and the max-pool backward
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
bash tools/r2_sanitizer.sh > gpurun_out/r2c8_sanitizer.log 2>&1
( timeout 600 ncu --set full --clock-control none --import-source on -k regex:"bn_bwd_(reduce|apply)_kernel" -s 380 -c 10 -f \
    -o gpurun_out/r2c8_bn_bwd python bench.py --steps 1 --warmup 3 --no-cpu-baseline ) > gpurun_out/r2c8_ncu_bn.log 2>&1
echo "ncu bn rc=$?" >> gpurun_out/r2c8_ncu_bn.log
( timeout 600 ncu --set full --clock-control none --import-source on -k regex:"maxpool_bwd_kernel|bn_apply_kernel" -s 147 -c 4 -f \
    -o gpurun_out/r2c8_pool python bench.py --steps 1 --warmup 3 --no-cpu-baseline ) > gpurun_out/r2c8_ncu_pool.log 2>&1
echo "ncu pool rc=$?" >> gpurun_out/r2c8_ncu_pool.log
tail -20 gpurun_out/r2c8_sanitizer.log; tail -2 gpurun_out/r2c8_ncu_bn.log; tail -2 gpurun_out/r2c8_ncu_pool.log; ls -la gpurun_out/r2c8_*.ncu-rep
exit 0
