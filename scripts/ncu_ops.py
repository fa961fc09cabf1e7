"""Launches the attention / GroupNorm / LayerNorm kernels once each at the UNet's N=16 shapes (for ncu captures)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from reface_b200.runtime import Engine
eng = Engine(0, arena_bytes=8 << 30)
for k, v in os.environ.items():
    if k.startswith("RFB_") and k != "RFB_CPU_THREADS":
        eng.set_option(k[4:].lower(), int(v))
what = os.environ.get("WHAT", "attn,gn,ln").split(",")
if "attn" in what:
    N, L, heads, d = 16, 4096, 8, 40
    qkv = torch.randn(N, L, 3 * heads * d, device="cuda").half().float()
    eng.op_attention(qkv, heads)
if "gn" in what:
    eng.bench_norm(0, 16, 320, 64, 64, iters=1)
    eng.bench_norm(0, 8, 128, 512, 512, iters=1)
if "ln" in what:
    eng.bench_norm(1, 16, 320, 64, 64, iters=1)
torch.cuda.synchronize()
print("done")
