"""Tile-shape sweep for the text / visual encoder GEMMs (M = batch * words = 1536 rows): GPU time per launch of each
(tile width, single CTA / CTA pair) variant, measured by replaying a CUDA graph of 20 back-to-back launches (Python
launch overhead excluded).  Run on the B200 box via gpurun; not a pytest file."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from univl_b200 import lib  # noqa: E402

REPS = 20


def time_variant(a_mn, b_mn, M, N, K, epi, bn_code):
    A = (torch.randn(K, M, device="cuda") if a_mn else torch.randn(M, K, device="cuda")).bfloat16()
    B = (torch.randn(K, N, device="cuda") if b_mn else torch.randn(N, K, device="cuda")).bfloat16()
    f32 = epi in (4, 5)
    out = torch.zeros(M, N, device="cuda", dtype=torch.float32 if f32 else torch.bfloat16)
    bias = torch.randn(N, device="cuda")
    aux_in = torch.randn(M, N, device="cuda").bfloat16()
    aux_out = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16)

    def go(stream):
        lib.call("univl_gemm_bf16", A.data_ptr(), A.stride(0), int(a_mn), B.data_ptr(), B.stride(0), int(b_mn), M, N,
                 K, out.data_ptr(), N, epi, bias.data_ptr(), aux_in.data_ptr(), N, aux_out.data_ptr(), N, 1.0,
                 bn_code, 0, stream)

    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        go(side.cuda_stream)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        st = torch.cuda.current_stream().cuda_stream
        for _ in range(REPS):
            go(st)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(3):
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1000.0 / REPS)
    return best


def main():
    rows = int(os.environ.get("SWEEP_ROWS", "1536"))
    cases = [
        ("fwd qkv", 0, 0, rows, 2304, 768, 0), ("fwd out", 0, 0, rows, 768, 768, 0),
        ("fwd ffn1+gelu", 0, 0, rows, 3072, 768, 1), ("fwd ffn2", 0, 0, rows, 768, 3072, 0),
        ("dgrad qkv", 0, 1, rows, 768, 2304, 0), ("dgrad out", 0, 1, rows, 768, 768, 0),
        ("dgrad ffn2+gelu'", 0, 1, rows, 3072, 768, 2), ("dgrad ffn1", 0, 1, rows, 768, 3072, 0),
        ("wgrad qkv", 1, 1, 2304, 768, rows, 5), ("wgrad out", 1, 1, 768, 768, rows, 5),
        ("wgrad ffn2", 1, 1, 768, 3072, rows, 5), ("wgrad ffn1", 1, 1, 3072, 768, rows, 5),
    ]
    variants = [("auto", 0), ("bn64", 64), ("bn128", 128), ("bn256", 256 + 1024), ("pair256", 256 + 512)]
    for name, a_mn, b_mn, M, N, K, epi in cases:
        res = {}
        for vname, code in variants:
            res[vname] = round(time_variant(a_mn, b_mn, M, N, K, epi, code), 2)
        gf = 2.0 * M * N * K / 1e9
        print("SWEEP " + json.dumps(dict(case=name, M=M, N=N, K=K, epi=epi, us=res,
                                         best_tflops=round(gf / min(res.values()) / 1e3, 1))), flush=True)


if __name__ == "__main__":
    main()
