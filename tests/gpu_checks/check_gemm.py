"""GPU bring-up check for the tcgen05 GEMM (run on the B200 box via gpurun; not a pytest file).

Compares univl_gemm_bf16 against torch fp32 matmul of the same bf16-rounded operands for one operand-major
combination per process (a trapped kernel poisons the CUDA context, so combos are isolated by the caller).
"""
import argparse
import ctypes
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from univl_b200 import lib  # noqa: E402


def run_case(a_mn, b_mn, M, N, K, bn, epi, split_k=0, verbose=False):
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N * 3 + K)
    A = (torch.randn(M, K, device="cuda", generator=g) * 0.5).bfloat16()   # logical A[m,k]
    B = (torch.randn(N, K, device="cuda", generator=g) * 0.5).bfloat16()   # logical B[n,k]
    A_mem = A.t().contiguous() if a_mn else A                               # MN-major: stored [K, M]
    B_mem = B.t().contiguous() if b_mn else B
    lda = A_mem.stride(0)
    ldb = B_mem.stride(0)
    ref = A.float() @ B.float().t()
    bias = torch.randn(N, device="cuda", generator=g)
    aux_in = torch.randn(M, N, device="cuda", generator=g).bfloat16()
    aux_out = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16)
    alpha = 1.0
    if epi in (4, 5):
        out = torch.zeros(M, N, device="cuda", dtype=torch.float32)
        if epi == 5:
            out.fill_(1.0)
    else:
        out = torch.full((M, N), 7.0, device="cuda", dtype=torch.bfloat16)
    stream = torch.cuda.current_stream().cuda_stream
    lib.call("univl_gemm_bf16", A_mem.data_ptr(), lda, int(a_mn), B_mem.data_ptr(), ldb, int(b_mn), M, N, K,
             out.data_ptr(), out.stride(0), epi, bias.data_ptr(), aux_in.data_ptr(), aux_in.stride(0),
             aux_out.data_ptr(), aux_out.stride(0), alpha, bn, split_k, stream)
    torch.cuda.synchronize()
    if epi == 0:
        exp = ref + bias
    elif epi == 1:
        pre = ref + bias
        exp = torch.nn.functional.gelu(pre)
    elif epi == 2:
        x = aux_in.float()
        cdf = 0.5 * (1 + torch.erf(x / 2 ** 0.5))
        pdf = torch.exp(-0.5 * x * x) / (2 * 3.141592653589793) ** 0.5
        exp = ref * (cdf + x * pdf)
    elif epi == 3:
        exp = ref + aux_in.float()
    elif epi == 4:
        exp = ref + bias
    else:
        exp = ref + 1.0
    got = out.float()
    err = (got - exp).abs().max().item()
    scale = exp.abs().max().item() + 1e-6
    tol = 2e-2 * scale if epi not in (4, 5) else 2e-3 * scale
    ok = err <= tol
    if epi == 1:
        err2 = (aux_out.float() - (ref + bias)).abs().max().item()
        ok = ok and err2 <= 2e-2 * scale
    if (not ok) and verbose:
        bad = ((got - exp).abs() > tol).nonzero()
        print("   first bad idx:", bad[:8].tolist(), "n_bad", bad.shape[0], "of", M * N)
        print("   got[0,:8]", got[0, :8].tolist())
        print("   exp[0,:8]", exp[0, :8].tolist())
    return ok, err, scale


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--combo", default="KK")
    ap.add_argument("--perf", action="store_true")
    args = ap.parse_args()
    a_mn = args.combo[0] == "M"
    b_mn = args.combo[1] == "M"
    results = []
    shapes = [(128, 64, 64), (128, 128, 128), (256, 256, 192), (300, 200, 136), (1536, 768, 768),
              (1536, 3072, 768), (1536, 768, 3072), (520, 30522, 768)]
    all_ok = True
    for (M, N, K) in shapes:
        for bn in (64, 128, 256):
            if N < bn and bn > 64:
                continue
            if N > 4096 and bn != 256:
                continue
            for epi in ((0, 4, 5) if (M, N, K) != (1536, 768, 768) else (0, 1, 2, 3, 4, 5)):
                if a_mn or b_mn:
                    # MN-major storage needs ld % 8 == 0 on the MN extent
                    if (a_mn and M % 8) or (b_mn and N % 8):
                        continue
                ok, err, scale = run_case(a_mn, b_mn, M, N, K, bn, epi, verbose=True)
                all_ok &= ok
                line = dict(combo=args.combo, M=M, N=N, K=K, bn=bn, epi=epi, ok=bool(ok), err=err, scale=scale)
                results.append(line)
                print(("PASS " if ok else "FAIL ") + json.dumps(line), flush=True)
    if args.perf and all_ok:
        for (M, N, K, bn) in [(98304, 768, 768, 256), (98304, 3072, 768, 256), (98304, 768, 3072, 256),
                              (98304, 2304, 768, 256), (98304, 768, 768, 128), (1536, 768, 768, 64),
                              (1536, 3072, 768, 128), (8192, 8192, 8192, 256)]:
            A = torch.randn(K, M, device="cuda").bfloat16() if a_mn else torch.randn(M, K, device="cuda").bfloat16()
            B = torch.randn(K, N, device="cuda").bfloat16() if b_mn else torch.randn(N, K, device="cuda").bfloat16()
            out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
            st = torch.cuda.current_stream().cuda_stream

            def go():
                lib.call("univl_gemm_bf16", A.data_ptr(), A.stride(0), int(a_mn), B.data_ptr(), B.stride(0),
                         int(b_mn), M, N, K, out.data_ptr(), N, 0, None, None, 0, None, 0, 1.0, bn, 0, st)
            for _ in range(3):
                go()
            torch.cuda.synchronize()
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                go()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            t0 = time.time()
            for _ in range(3):
                torch.matmul(A.t() if a_mn else A, B if b_mn else B.t())
            torch.cuda.synchronize()
            e0.record()
            for _ in range(10):
                torch.matmul(A.t() if a_mn else A, B if b_mn else B.t())
            e1.record()
            torch.cuda.synchronize()
            ms_t = e0.elapsed_time(e1) / 10
            tf = 2.0 * M * N * K / ms / 1e9
            print("PERF " + json.dumps(dict(combo=args.combo, M=M, N=N, K=K, bn=bn, ms=ms, tflops=tf,
                                            cublas_ms=ms_t, cublas_tflops=2.0 * M * N * K / ms_t / 1e9)), flush=True)
    print("SUMMARY combo=%s all_ok=%s n=%d" % (args.combo, all_ok, len(results)), flush=True)
    sys.exit(0 if all_ok else 1)


if __name__ == "__main__":
    main()
