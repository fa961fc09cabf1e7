"""First-contact diagnostics for the tcgen05 GEMM on a real B200 (run under gpurun)."""
import math
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from reface_b200.runtime import Engine

torch.manual_seed(0)
eng = Engine(0, arena_bytes=8 << 30)
print("device", torch.cuda.get_device_name(0), flush=True)


def report(name, y, ref):
    err = (y - ref).abs()
    print(f"{name}: max_err={float(err.max()):.4e} ref_max={float(ref.abs().max()):.3e} "
          f"bad={(err > 1e-2 * ref.abs().max()).float().mean().item():.4f}", flush=True)
    return float(err.max()) / float(ref.abs().max())


# 1. identity weight: y must equal x (reveals swizzle / descriptor problems)
for (M, K, N) in [(128, 64, 64), (128, 128, 128), (256, 320, 320)]:
    x = torch.randn(M, K, device="cuda").half().float()
    w = torch.zeros(N, K, device="cuda")
    w[torch.arange(min(N, K)), torch.arange(min(N, K))] = 1.0
    y = eng.op_linear(x, w)
    torch.cuda.synchronize()
    r = report(f"identity {M}x{K}x{N}", y, F.linear(x, w))
    if r > 1e-2:
        print(" x[0,:8]", x[0, :8].tolist())
        print(" y[0,:8]", y[0, :8].tolist())
        print(" y[1,:8]", y[1, :8].tolist())
        print(" y[0,8:16]", y[0, 8:16].tolist())
# 2. random
for (M, K, N) in [(128, 64, 32), (128, 256, 256), (1000, 768, 1280), (4096, 2880, 320)]:
    x = torch.randn(M, K, device="cuda").half().float()
    w = (torch.randn(N, K, device="cuda") / math.sqrt(K)).half().float()
    b = torch.randn(N, device="cuda")
    y = eng.op_linear(x, w, b)
    torch.cuda.synchronize()
    report(f"random {M}x{K}x{N}", y, F.linear(x, w, b))
# 3. conv via TMA halo boxes
for (N, C, H, O) in [(1, 64, 8, 64), (2, 64, 16, 64), (2, 320, 64, 320), (1, 128, 128, 128)]:
    x = torch.randn(N, C, H, H, device="cuda").half().float()
    w = (torch.randn(O, C, 3, 3, device="cuda") / math.sqrt(9 * C)).half().float()
    b = torch.randn(O, device="cuda")
    y = eng.op_conv2d(x, w, b)
    torch.cuda.synchronize()
    ref = F.conv2d(x, w, b, padding=1)
    r = report(f"conv {N}x{C}x{H}->{O}", y, ref)
    if r > 1e-2:
        e = (y - ref).abs().amax(dim=(0, 1))
        print(" err map rows:", e.amax(dim=1)[:8].tolist(), " cols:", e.amax(dim=0)[:8].tolist())
# 4. timing of a big GEMM
x = torch.randn(65536, 2880, device="cuda").half().float()
w = (torch.randn(320, 2880, device="cuda") / 50).half().float()
for bn, st in [(0, 0)]:
    y = eng.op_linear(x, w)
torch.cuda.synchronize()
print("launches", eng.launch_count)
print("DIAG DONE")
