#!/bin/bash
# round-2 GPU call 14: ncu launch list (time + DRAM bytes) of one step of the final build, and ncu --set full of the
# FINAL conv kernels on single batch-256 layers + the BN backward kernels inside a step
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
( timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1400 --csv \
    --log-file gpurun_out/launches_r2f.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline ) > gpurun_out/r2c14_ncu.log 2>&1
prof() {  # name, layer substring, form
  ( timeout 300 ncu --set full --clock-control none --import-source on -k regex:igemm_kernel -s 2 -c 1 -f \
      -o gpurun_out/r2f_$1 python tests/cta2_check.py one "$2" $3 ) > gpurun_out/r2c14_ncu_$1.log 2>&1
  echo "ncu $1 rc=$?" >> gpurun_out/r2c14_ncu_$1.log
}
prof l1c2_fprop "l1.c2 " fprop
prof l3c2_fprop "l3.c2 " fprop
prof l3c3_fprop "l3.c3 " fprop
prof l1c3_fprop "l1.c3 " fprop
prof l2c2_fprop "l2.c2 " fprop
prof l3c1_wgrad "l3.c1 " wgrad
prof l1c2_wgrad "l1.c2 " wgrad
ls -la gpurun_out/*.ncu-rep; tail -2 gpurun_out/r2c14_ncu.log
exit 0
