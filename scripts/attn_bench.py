"""Times the fused attention kernel alone (CUDA events around the launch) at the UNet's shapes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from reface_b200.runtime import Engine
eng = Engine(0, arena_bytes=12 << 30)
for (N, L, heads, d) in [(16, 4096, 8, 40), (16, 1024, 8, 80)]:
    C = heads * d
    qkv = torch.randn(N, L, 3 * C, device="cuda").half().float()
    y = eng.op_attention(qkv, heads)
    q, k, v = qkv.chunk(3, dim=-1)
    sp = lambda t: t.reshape(N, L, heads, d).transpose(1, 2)
    ref = F.scaled_dot_product_attention(sp(q), sp(k), sp(v)).transpose(1, 2).reshape(N, L, C)
    err = float((y - ref).abs().max())
    best = 1e9
    for _ in range(3):
        eng.set_option("profile", 1)
        eng.op_attention(qkv, heads)
        ms, fl, n = eng.profile_read()
        eng.set_option("profile", 0)
        best = min(best, ms)
    print(f"flash N={N} L={L} heads={heads} d={d}: {best*1e3:8.1f} us  {4.0*L*L*d*N*heads/best/1e9:7.1f} TFLOP/s (algorithmic)  max_err={err:.2e}", flush=True)
