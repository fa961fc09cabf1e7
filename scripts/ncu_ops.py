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
if "gn3" in what:   # GroupNorm + SiLU fed by producer statistics: gn_finalize3 + gn_apply3 (streaming pass)
    eng.bench_norm(2, 16, 320, 64, 64, iters=1)
if "ln" in what:
    eng.bench_norm(1, 16, 320, 64, 64, iters=1)
if "lin" in what:   # the HBM-bound K=320 projection of the 64x64 level with bias + residual (attn1.to_out / proj_out)
    M, K, N = 65536, 320, 320
    x = torch.randn(M, K, device="cuda").half().float(); w = torch.randn(N, K, device="cuda").half().float() / 18
    b = torch.randn(N, device="cuda"); r = torch.randn(M, N, device="cuda").half().float()
    eng.op_linear(x, w, b, residual=r)
if "conv" in what:  # the dominant long-K implicit-GEMM conv (ResBlock 3x3, 640 ch at 32x32, N=16)
    x = torch.randn(16, 640, 32, 32, device="cuda").half().float(); w = torch.randn(640, 640, 3, 3, device="cuda").half().float() / 76
    eng.op_conv2d(x, w, torch.randn(640, device="cuda"))
torch.cuda.synchronize()
print("done")
