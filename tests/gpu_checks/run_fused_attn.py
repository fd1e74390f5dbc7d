"""Timing / profiling driver for the fused attention kernels at BASELINE shapes.

    python tests/gpu_checks/run_fused_attn.py --n_seq 1024 --S 96 --iters 20      # CUDA-event timings, fused vs unfused
    ncu --set full -k regex:fused_ ... python tests/gpu_checks/run_fused_attn.py --n_seq 1024 --S 96 --iters 1 --only fused
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from univl_b200 import ops  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n_seq", type=int, default=1024)
    ap.add_argument("--S", type=int, default=96)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--p", type=float, default=0.1)
    ap.add_argument("--only", default="both", choices=["both", "fused", "unfused"])
    a = ap.parse_args()
    H, dev = 768, "cuda"
    n_seq, S = a.n_seq, a.S
    T = n_seq * S
    g = torch.Generator(device=dev).manual_seed(1)
    x = torch.randn(T, H, device=dev, generator=g).bfloat16()
    w = (torch.randn(3 * H, H, device=dev, generator=g) * 0.04).bfloat16()
    b = torch.randn(3 * H, device=dev, generator=g) * 0.2
    lens = torch.randint(S // 2, S + 1, (n_seq,), generator=torch.Generator().manual_seed(1)).to(dev)
    mask = (torch.arange(S, device=dev).unsqueeze(0) < lens.unsqueeze(1)).long()
    spec = ops.MaskSpec(mask)
    rng = torch.tensor([123, 0], dtype=torch.int64, device=dev)
    d_o = torch.randn(T, H, device=dev, generator=g).bfloat16()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > L2 (126 MB)

    def timed(fn, n):
        ts = []
        for _ in range(n):
            flush.zero_()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            fn()
            e.record()
            torch.cuda.synchronize()
            ts.append(s.elapsed_time(e))
        ts.sort()
        return ts[len(ts) // 2]

    qkv_flops = 2.0 * T * 3 * H * H
    core_flops = 4.0 * T * S * H
    out = {}
    if a.only in ("both", "fused"):
        o, lse, qkv = ops.fused_qkv_attention_fwd(x, w, b, n_seq, S, spec, p=a.p, seed=rng.data_ptr(), stream=3)
        dqkv = torch.empty_like(qkv)
        dbias = torch.zeros(3 * H, device=dev)
        f = lambda: ops.fused_qkv_attention_fwd(x, w, b, n_seq, S, spec, p=a.p, seed=rng.data_ptr(), stream=3)
        f2 = lambda: ops.fused_qkv_attention_fwd(x, w, b, n_seq, S, spec, p=a.p, seed=rng.data_ptr(), stream=3,
                                                 save_qkv=False)
        fb = lambda: ops.fused_attention_bwd(qkv, o, lse, d_o, dqkv, n_seq, S, spec, p=a.p, seed=rng.data_ptr(),
                                             stream=3, dbias=dbias)
        for _ in range(2):
            f(); f2(); fb()
        torch.cuda.synchronize()
        if a.iters > 1:
            out["fused_fwd_ms"] = timed(f, a.iters)
            out["fused_fwd_noqkv_ms"] = timed(f2, a.iters)
            out["fused_bwd_ms"] = timed(fb, a.iters)
            out["fused_fwd_tflops"] = (qkv_flops + core_flops) / out["fused_fwd_ms"] * 1e-9
    if a.only in ("both", "unfused"):
        qkv = torch.empty(T, 3 * H, dtype=torch.bfloat16, device=dev)

        def u():
            ops.gemm(x, w, T, 3 * H, H, qkv, bias=b)
            return ops.attention_fwd(qkv[:, :H], qkv[:, H:2 * H], qkv[:, 2 * H:], n_seq, S, S, spec, p=a.p,
                                     seed=rng.data_ptr(), stream=3)
        o, lse = u()
        dq2 = torch.empty_like(qkv)
        db = torch.zeros(3, H, device=dev)
        ub = lambda: ops.attention_bwd(qkv[:, :H], qkv[:, H:2 * H], qkv[:, 2 * H:], o, lse, d_o, dq2[:, :H],
                                       dq2[:, H:2 * H], dq2[:, 2 * H:], n_seq, S, S, spec, p=a.p, seed=rng.data_ptr(),
                                       stream=3, dbias=(db[0], db[1], db[2]))
        for _ in range(2):
            u(); ub()
        torch.cuda.synchronize()
        if a.iters > 1:
            out["unfused_fwd_ms"] = timed(u, a.iters)
            out["unfused_bwd_ms"] = timed(ub, a.iters)
    out.update(n_seq=n_seq, S=S, T=T, p=a.p, qkv_gflop=qkv_flops * 1e-9, core_gflop=core_flops * 1e-9)
    print(out)


if __name__ == "__main__":
    main()
