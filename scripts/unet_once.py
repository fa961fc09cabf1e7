"""Times UNet forwards (N=16, L=64 by default) with synthetic weights; used under ncu for the launch list."""
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from reface_b200 import synth
from reface_b200.runtime import Engine

N = int(os.environ.get("N", 16)); L = int(os.environ.get("L", 64)); REPS = int(os.environ.get("REPS", 5))
dev = torch.device("cuda", 0)
flat = synth.random_flat(dev, 0)
sd = {k: v for k, v in synth.state_dict_from_flat(flat).items() if k.startswith("model.diffusion_model.")}
eng = Engine(0)
for k, v in os.environ.items():
    if k.startswith("RFB_"):
        eng.set_option(k[4:].lower(), int(v))
eng.load_state_dict(sd)
eng.build_unet()
x = torch.randn(N, 9, L, L, device=dev); t = torch.full((N,), 981, device=dev, dtype=torch.long)
ctx = torch.randn(N, 1, 768, device=dev)
l0 = eng.launch_count
eps = eng.unet_forward(x, t, ctx); torch.cuda.synchronize()
print("launches per forward", eng.launch_count - l0, "finite", bool(torch.isfinite(eps).all()), "absmax", float(eps.abs().max()))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(REPS):
    eng.unet_forward(x, t, ctx)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / REPS
print(f"unet forward N={N} L={L}: {ms:.3f} ms  -> {12.75e12 * N / 16 / (ms / 1e3) / 1e12 if L == 64 else 0:.1f} TFLOP/s ref-equivalent")
eng.set_option("profile", 1)
eng.unet_forward(x, t, ctx)
gms, gfl, gn = eng.profile_read()
eng.set_option("profile", 0)
print(f"gemm_tc launches={gn} time={gms:.3f} ms alg_flops={gfl/1e12:.3f} TF -> {gfl/gms/1e9:.1f} TFLOP/s; arena peak {eng.arena_peak/2**30:.2f} GiB")
