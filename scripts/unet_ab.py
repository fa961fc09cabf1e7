"""A/B of engine options over the whole UNet forward (N=16, L=64 by default): builds the UNet once, then times
REPS forwards per option set.  Usage: python scripts/unet_ab.py "attn_flash=1" "gn_fused=0" "attn_flash=1,gn_fused=0,ln_vec=0"
(the empty set = current defaults is always run first and last)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from reface_b200 import synth
from reface_b200.runtime import Engine

N = int(os.environ.get("N", 16)); L = int(os.environ.get("L", 64)); REPS = int(os.environ.get("REPS", 5))
DEFAULTS = {"attn_flash": 4, "attn_poly": 0, "attn_stagger": 0, "gn_fused": 1, "gn_cluster": 16, "gn_threads": 512, "ln_vec": 1,
            "gemm_pair": 0, "gemm_wave_bn": 1, "gn_epi_stats": 1, "gemm_splitk": 1, "attn_pingpong": 0, "gn_fold": 1, "gemm_mcast": 1, "gemm_lean": 1, "pdl": 1, "gemm_mcast_big": 2}
dev = torch.device("cuda", 0)
flat = synth.random_flat(dev, 0)
sd = {k: v for k, v in synth.state_dict_from_flat(flat).items() if k.startswith("model.diffusion_model.")}
eng = Engine(0)
eng.load_state_dict(sd)
eng.build_unet()
x = torch.randn(N, 9, L, L, device=dev); t = torch.full((N,), 981, device=dev, dtype=torch.long)
ctx = torch.randn(N, 1, 768, device=dev)
ROUNDS = int(os.environ.get("ROUNDS", 3))
specs = [""] + sys.argv[1:]
base = None
res = {sp: [] for sp in specs}
tc = {sp: [] for sp in specs}
diff = {}


def apply(spec):
    opts = dict(DEFAULTS)
    for kv in filter(None, spec.split(",")):
        k, v = kv.split("=")
        opts[k] = int(v)
    for k, v in opts.items():
        eng.set_option(k, v)


# interleaved rounds (A B C A B C ...): the box's power-capped clocks drift by a few percent over seconds, a single
# back-to-back comparison is not reliable at that level; report the median per option set
for r in range(ROUNDS):
    for spec in specs:
        apply(spec)
        eps = eng.unet_forward(x, t, ctx); torch.cuda.synchronize()
        if base is None:
            base = eps.clone()
        diff[spec] = float((eps - base).abs().max() / base.abs().max())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(REPS):
            eng.unet_forward(x, t, ctx)
        e1.record(); torch.cuda.synchronize()
        res[spec].append(e0.elapsed_time(e1) / REPS)
        eng.set_option("profile", 1)
        eng.unet_forward(x, t, ctx)
        gms, gfl, gn = eng.profile_read()
        eng.set_option("profile", 0)
        tc[spec].append(gms)
med = lambda v: sorted(v)[len(v) // 2]
for spec in specs:
    print(f"[{spec or 'defaults':44s}] unet forward N={N} L={L}: median {med(res[spec]):7.3f} ms (runs " +
          " ".join(f"{v:.2f}" for v in res[spec]) + f")  tensor-core {med(tc[spec]):7.3f} ms  other {med(res[spec]) - med(tc[spec]):6.3f} ms  "
          f"diff vs defaults {diff[spec]:.2e}", flush=True)
