#!/bin/bash
# GPU bring-up of the CTA-pair GEMM: correctness of all operand-major combos with the pair kernel forced on, then perf
mkdir -p gpurun_out
rm -f gpurun_out/gemm_pair.log
for c in KK KM MM MK; do
  UNIVL_GEMM_PAIR=2 timeout 200 python tests/gpu_checks/check_gemm.py --combo $c --perf >> gpurun_out/gemm_pair.log 2>&1
  echo "exit $c $?" >> gpurun_out/gemm_pair.log
done
grep -c PASS gpurun_out/gemm_pair.log
grep "FAIL\|exit\|rror\|timed out" gpurun_out/gemm_pair.log | head -30
grep PERF gpurun_out/gemm_pair.log | head -8
