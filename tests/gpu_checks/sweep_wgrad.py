"""Tile / split-K sweep of the weight-gradient GEMMs at the encoder shapes (contraction over T = 1536 tokens):
dW[N_out, K_in] += dy[T, N_out]^T x[T, K_in]   (A = dy MN-major, B = x MN-major, fp32 reduce-add epilogue)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from univl_b200 import ops

dev = "cuda"
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, n=12):
    ts = []
    for _ in range(n):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


T = int(sys.argv[1]) if len(sys.argv) > 1 else 1536
for (No, Ki) in [(3072, 768), (768, 3072), (768, 768), (2304, 768)]:
    dy = torch.randn(T, No, device=dev).bfloat16()
    x = torch.randn(T, Ki, device=dev).bfloat16()
    dw = torch.zeros(No, Ki, device=dev)
    ref = dy.float().t() @ x.float()
    print("== dW[%d,%d] over T=%d  (%.1f GFLOP)" % (No, Ki, T, 2e-9 * No * Ki * T))
    for bn_name, bn in (("pair256", 256 + 512), ("single256", 256 + 1024), ("single128", 128), ("single64", 64)):
        for sk in (0, 1, 2, 3, 4, 6, 8, 12):
            try:
                dw.zero_()
                ops.gemm(dy, x, No, Ki, T, dw, epi=ops.EPI_ATOMIC, a_mn=True, b_mn=True, block_n=bn, split_k=sk)
                torch.cuda.synchronize()
                err = float((dw - ref).abs().max() / ref.abs().max())
                us = timed(lambda: ops.gemm(dy, x, No, Ki, T, dw, epi=ops.EPI_ATOMIC, a_mn=True, b_mn=True,
                                            block_n=bn, split_k=sk))
                print("  %-9s split %2d : %6.1f us  %6.0f TF/s  err %.1e" % (bn_name, sk, us, 2e-6 * No * Ki * T / us, err))
            except Exception as ex:  # noqa: BLE001
                print("  %-9s split %2d : failed %s" % (bn_name, sk, str(ex)[:80]))
