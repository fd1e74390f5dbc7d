"""Bring-up check of the fused attention kernels: every shape in its own process (a device fault is sticky), with
CUDA_LAUNCH_BLOCKING, printing error figures.   python tests/gpu_checks/check_fused_attn.py [fwd|bwd|all]"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
SHAPES = [(2, 96, 0, 0.0), (3, 48, 0, 0.0), (2, 128, 1, 0.0), (9, 16, 1, 0.0), (40, 96, 0, 0.0), (3, 48, 0, 0.25)]


def one(kind, n_seq, S, causal, p):
    sys.path.insert(0, ROOT)
    import torch
    from univl_b200 import ops
    H, h = 768, 12
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(S + n_seq)
    x = (torch.randn(n_seq * S, H, device=dev, generator=g)).bfloat16()
    w = (torch.randn(3 * H, H, device=dev, generator=g) * 0.04).bfloat16()
    b = torch.randn(3 * H, device=dev, generator=g) * 0.2
    lens = torch.randint(1, S + 1, (n_seq,), generator=torch.Generator().manual_seed(1)).to(dev)
    mask = (torch.arange(S, device=dev).unsqueeze(0) < lens.unsqueeze(1)).long()
    spec = ops.MaskSpec(mask, causal=bool(causal))
    rng = torch.tensor([123, 0], dtype=torch.int64, device=dev)
    o, lse, qkv = ops.fused_qkv_attention_fwd(x, w, b, n_seq, S, spec, p=p, seed=rng.data_ptr(), stream=3)
    torch.cuda.synchronize()
    qkv_ref = (x.float() @ w.float().t() + b).bfloat16()
    print("  qkv err %.4g" % float((qkv.float() - qkv_ref.float()).abs().max()), flush=True)
    if p == 0.0:
        o2, lse2 = ops.attention_fwd(qkv[:, :H], qkv[:, H:2 * H], qkv[:, 2 * H:], n_seq, S, S, spec)
        torch.cuda.synchronize()
        print("  ctx err vs unfused %.4g   lse err %.4g   finite %s" % (
            float((o.float() - o2.float()).abs().max()), float((lse - lse2).abs().max()),
            bool(torch.isfinite(o.float()).all())), flush=True)
    if kind == "fwd":
        return
    d_o = torch.randn(n_seq * S, H, device=dev, generator=g).bfloat16()
    dqkv = torch.empty_like(qkv)
    dbias = torch.zeros(3 * H, device=dev)
    ops.fused_attention_bwd(qkv, o, lse, d_o, dqkv, n_seq, S, spec, p=p, seed=rng.data_ptr(), stream=3, dbias=dbias)
    torch.cuda.synchronize()
    dq2 = torch.empty_like(qkv)
    ops.attention_bwd(qkv[:, :H], qkv[:, H:2 * H], qkv[:, 2 * H:], o, lse, d_o, dq2[:, :H], dq2[:, H:2 * H],
                      dq2[:, 2 * H:], n_seq, S, S, spec, p=p, seed=rng.data_ptr(), stream=3, rng_layout=1)
    torch.cuda.synchronize()
    for n, nm in enumerate(("dq", "dk", "dv")):
        a, c = dqkv[:, n * H:(n + 1) * H].float(), dq2[:, n * H:(n + 1) * H].float()
        print("  %s max err %.4g (scale %.3g) rel fro %.4g" % (nm, float((a - c).abs().max()), float(c.abs().max()),
                                                                float((a - c).norm() / c.norm())), flush=True)
    print("  dbias err %.4g" % float((dbias - dqkv.float().sum(0)).abs().max()), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 2:
        one(sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), float(sys.argv[5]))
        sys.exit(0)
    kind = sys.argv[1] if len(sys.argv) > 1 else "all"
    env = dict(os.environ, CUDA_LAUNCH_BLOCKING="1")
    for sh in SHAPES:
        print("== %s n_seq=%d S=%d causal=%d p=%g" % ((kind,) + sh), flush=True)
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "bwd" if kind != "fwd" else "fwd"] +
                           [str(v) for v in sh], env=env, capture_output=True, text=True, timeout=120)
        print(r.stdout[-1500:], flush=True)
        if r.returncode != 0:
            print("  FAILED rc=%d\n%s" % (r.returncode, r.stderr[-1200:]), flush=True)
