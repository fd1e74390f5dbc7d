#!/bin/bash
# GPU bring-up of the persistent GEMMs: all operand-major combos with (a) 1-CTA kernels only, (b) the CTA-pair kernel
# forced on for every BLOCK_N = 256 case, (c) the manual-store fallback epilogue
mkdir -p gpurun_out
rm -f gpurun_out/gemm_all.log
for mode in "UNIVL_GEMM_PAIR=1" "UNIVL_GEMM_PAIR=2" "UNIVL_GEMM_MANUAL_EPILOGUE=1"; do
  for c in KK KM MM MK; do
    perf=""; [ "$c" = "KK" ] && perf="--perf"
    env $mode timeout 200 python tests/gpu_checks/check_gemm.py --combo $c $perf >> gpurun_out/gemm_all.log 2>&1
    echo "exit $mode $c $?" >> gpurun_out/gemm_all.log
  done
done
grep -c PASS gpurun_out/gemm_all.log
grep "FAIL\|exit\|rror\|timed out" gpurun_out/gemm_all.log | head -40
grep PERF gpurun_out/gemm_all.log | cut -c1-200
