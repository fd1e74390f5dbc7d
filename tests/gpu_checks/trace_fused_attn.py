"""Per-phase timeline of the fused forward kernel (CTA 0, first items) from clock64 stamps — bring-up aid."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from univl_b200 import lib, ops

n_seq, S, H, dev = 1024, int(sys.argv[1]) if len(sys.argv) > 1 else 96, 768, "cuda"
g = torch.Generator(device=dev).manual_seed(1)
x = torch.randn(n_seq * S, H, device=dev, generator=g).bfloat16()
w = (torch.randn(3 * H, H, device=dev, generator=g) * 0.04).bfloat16()
b = torch.randn(3 * H, device=dev, generator=g) * 0.2
mask = torch.ones(n_seq, S, dtype=torch.long, device=dev)
rng = torch.tensor([1, 0], dtype=torch.int64, device=dev)
spec = ops.MaskSpec(mask)
for _ in range(2):
    ops.fused_qkv_attention_fwd(x, w, b, n_seq, S, spec, p=0.1, seed=rng.data_ptr(), stream=3)
trace = torch.zeros(16 * 16, dtype=torch.int64, device=dev)
L = lib.load()
L.univl_debug_set_fused_attention_trace.argtypes = [ctypes.c_void_p]
L.univl_debug_set_fused_attention_trace(trace.data_ptr())
ops.fused_qkv_attention_fwd(x, w, b, n_seq, S, spec, p=0.1, seed=rng.data_ptr(), stream=3)
torch.cuda.synchronize()
L.univl_debug_set_fused_attention_trace(None)
t = trace.view(16, 16).cpu()
t0 = int(t[0, 0])
names = ["proj_start", "proj_end", "acc_full_seen", "qkv_drained", "S_issue", "s_full_seen", "p_written", "PV_issue",
         "pv_done_seen", "O_drained", "store_read_done"]
print("cycles relative to proj_start of item 0 (CTA 0)")
print("item " + " ".join("%15s" % n for n in names))
for j in range(10):
    print("%4d " % j + " ".join("%15d" % (int(t[j, e]) - t0 if int(t[j, e]) else -1) for e in range(len(names))))
