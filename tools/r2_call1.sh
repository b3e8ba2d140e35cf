#!/bin/bash
# round-2 GPU call 1: (a) the never-run CTA-pair variants (parity, then per-layer timing incl. wgrad),
# (b) ncu --set full of the shipped conv kernels on single layers (stall reasons of the TMA / MMA / epilogue warps)
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
run_mode() {  # tag, env...
  tag=$1; shift
  ( env "$@" timeout 300 python tests/cta2_check.py parity ) > gpurun_out/r2c1_parity_$tag.log 2>&1
  rc=$?
  echo "parity rc=$rc" >> gpurun_out/r2c1_parity_$tag.log
  if [ $rc -eq 0 ]; then
    ( env "$@" timeout 300 python tests/cta2_check.py time ) > gpurun_out/r2c1_time_$tag.log 2>&1
  fi
}
( timeout 300 python tests/cta2_check.py time ) > gpurun_out/r2c1_time_default.log 2>&1
run_mode cta2_2 DIRB200_CTA2=2
run_mode cta2_2i DIRB200_CTA2=2 DIRB200_IM2COL=1
run_mode cta2_3 DIRB200_CTA2=3
run_mode cta2_3i DIRB200_CTA2=3 DIRB200_IM2COL=1
prof() {  # name, layer substring, form
  ( timeout 400 ncu --set full --clock-control none --import-source on -k regex:igemm_kernel -s 2 -c 1 -f \
      -o gpurun_out/r2c1_$1 python tests/cta2_check.py one "$2" $3 ) > gpurun_out/r2c1_ncu_$1.log 2>&1
  echo "ncu $1 rc=$?" >> gpurun_out/r2c1_ncu_$1.log
}
prof l3c2_fprop "l3.c2 " fprop
prof l2c2_fprop "l2.c2 " fprop
prof l4c1_fprop "l4.c1 " fprop
prof l1c3_fprop "l1.c3 " fprop
prof l3c2_wgrad "l3.c2 " wgrad
prof l1c2_fprop "l1.c2 " fprop
for f in gpurun_out/r2c1_parity_*.log; do echo "== $f"; grep -E "FAIL|parity" $f | tail -4; done
for f in gpurun_out/r2c1_time_*.log; do echo "== $f"; tail -1 $f; done
ls -la gpurun_out/*.ncu-rep
exit 0
