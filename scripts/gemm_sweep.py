"""GEMM kernel sweep on a B200: TFLOP/s of the tensor-core kernel alone (CUDA events around the launch) for
conv-like shapes of the UNet, across tile width / 2-CTA pairing / k-merge / stage count."""
import math
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from reface_b200.runtime import Engine

eng = Engine(0, arena_bytes=12 << 30)
SHAPES = [(65536, 2880, 320), (16384, 5760, 640), (4096, 11520, 1280), (65536, 320, 960), (16384, 2560, 640)]
if len(sys.argv) > 1:
    SHAPES = [tuple(int(v) for v in a.split("x")) for a in sys.argv[1:]]


def run(M, K, N, **opts):
    base = dict(gemm_pair=1, gemm_kmerge=2, gemm_bn=0, gemm_stages=0, gemm_persistent=1)
    base.update(opts)
    for k, v in base.items():
        eng.set_option(k, v)
    x = torch.randn(M, K, device="cuda").half().float()
    w = (torch.randn(N, K, device="cuda") / math.sqrt(K)).half().float()
    b = torch.randn(N, device="cuda")
    y = eng.op_linear(x, w, b)          # warm
    torch.cuda.synchronize()
    err = float((y - F.linear(x, w, b)).abs().max())
    best = 1e9
    for _ in range(3):
        eng.set_option("profile", 1)
        eng.op_linear(x, w, b)
        ms, fl, n = eng.profile_read()
        eng.set_option("profile", 0)
        best = min(best, ms)
    return 2.0 * M * N * K / best / 1e9, err


for (M, K, N) in SHAPES:
    print(f"--- M={M} K={K} N={N}", flush=True)
    for name, o in [("pair km2 auto", {}), ("pair km1 auto", dict(gemm_kmerge=1)), ("single persist", dict(gemm_pair=0)),
                    ("gen1 non-persist", dict(gemm_pair=0, gemm_persistent=0)),
                    ("pair km2 bn128", dict(gemm_bn=128)), ("pair km2 bn256", dict(gemm_bn=256)),
                    ("pair km2 bn64", dict(gemm_bn=64)), ("pair km2 bn192", dict(gemm_bn=192)),
                    ("pair km2 st3", dict(gemm_stages=3)), ("pair km2 st2", dict(gemm_stages=2)),
                    ("single bn256", dict(gemm_pair=0, gemm_bn=256)), ("single bn128", dict(gemm_pair=0, gemm_bn=128))]:
        try:
            tf, err = run(M, K, N, **o)
            print(f"  {name:18s} {tf:8.1f} TFLOP/s  max_err={err:.2e}", flush=True)
        except Exception as e:
            print(f"  {name:18s} FAILED {str(e)[:100]}", flush=True)
print("SWEEP DONE")
