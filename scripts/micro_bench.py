"""Kernel-level A/B timings at the UNet's shapes (N=16 = CFG batch of 8 faces at 512x512):
attention (flash v1 / v2), GroupNorm (3-kernel / fused cluster) and LayerNorm (warp-per-row / vectorised).
All times are CUDA-event times of the kernels alone; HBM fractions use the algorithmic bytes (read + write fp16)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from reface_b200.runtime import Engine

peaks = {}
pp = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
if os.path.exists(pp):
    peaks = json.load(open(pp))
HBM = float(peaks.get("hbm_gbs", 6549.1))

eng = Engine(0, arena_bytes=16 << 30)
out = {"attention": [], "groupnorm": [], "layernorm": []}

for (N, L, heads, d) in ([] if os.environ.get("ONLY") == "gn" else [(16, 4096, 8, 40), (16, 1024, 8, 80), (4, 16384, 8, 40)]):
    C = heads * d
    qkv = torch.randn(N, L, 3 * C, device="cuda").half().float()
    q, k, v = qkv.chunk(3, dim=-1)
    sp = lambda t: t.reshape(N, L, heads, d).transpose(1, 2)
    ref = F.scaled_dot_product_attention(sp(q), sp(k), sp(v)).transpose(1, 2).reshape(N, L, C)
    for mode, poly, pad in ((4, 0, 1), (4, 1, 1), (4, 1, 0), (4, 0, 1), (4, 1, 1), (4, 1, 0), (4, 0, 0)):   # pad = MUFU turn-taking on/off
        eng.set_option("attn_flash", mode)
        eng.set_option("attn_poly", poly)
        eng.set_option("attn_pingpong", pad)
        y = eng.op_attention(qkv, heads)
        err = float((y - ref).abs().max())
        best = 1e9
        for _ in range(4):
            eng.set_option("profile", 1)
            eng.op_attention(qkv, heads)
            ms, fl, n = eng.profile_read()
            eng.set_option("profile", 0)
            best = min(best, ms)
        tf = 4.0 * L * L * d * N * heads / best / 1e9
        exps = N * heads * L * L / (best * 1e-3) / 1e12
        print(f"attention v{mode} poly={poly} pingpong={pad} N={N} L={L} d={d}: {best*1e3:8.1f} us  {tf:7.1f} TFLOP/s  {exps:5.2f} Texp/s  "
              f"max_err={err:.2e}", flush=True)
        out["attention"].append({"mode": mode, "poly": poly, "N": N, "L": L, "d": d, "us": best * 1e3, "tflops": tf, "err": err})
    del qkv, ref, y
eng.set_option("attn_flash", 4)
eng.set_option("attn_poly", 0)
eng.set_option("attn_pingpong", 0)
if os.environ.get("ONLY") == "attn":
    sys.exit(0)
if os.environ.get("ONLY") == "gn":
    pass

for (N, Cc, H) in [(16, 320, 64), (16, 640, 64), (16, 960, 64), (16, 640, 32), (16, 1280, 32), (16, 1920, 32),
                   (16, 1280, 16), (16, 2560, 16), (16, 1280, 8), (16, 2560, 8), (8, 128, 512), (8, 256, 256), (2, 320, 64)]:
    row = {"N": N, "C": Cc, "H": H}
    gb = 2.0 * N * Cc * H * H * 2 / 1e9
    txt = []
    for name, fused, cl, th in (("split2", 0, 8, 512), ("c16t512", 1, 16, 512)):
        eng.set_option("gn_fused", fused); eng.set_option("gn_cluster", cl); eng.set_option("gn_threads", th)
        eng.set_option("gn_fused_max_elems", 1 << 40)
        ms = eng.bench_norm(0, N, Cc, H, H)
        row[name] = ms * 1e3
        txt.append(f"{name} {ms*1e3:6.1f} us ({gb / (ms * 1e-3) / HBM:.2f})")
    for bps in (2, 4, 8):
        eng.set_option("gn_apply_bps", bps)
        ms = eng.bench_norm(2, N, Cc, H, H)
        row[f"epi_stats_bps{bps}"] = ms * 1e3
        txt.append(f"epi_stats/bps{bps} {ms*1e3:6.1f} us ({gb / (ms * 1e-3) / HBM:.2f})")
    eng.set_option("gn_apply_bps", 4)
    print(f"groupnorm N={N} C={Cc} H={H}: " + "  ".join(txt), flush=True)
    out["groupnorm"].append(row)
eng.set_option("gn_cluster", 16); eng.set_option("gn_threads", 512); eng.set_option("gn_fused_max_elems", 2621440); eng.set_option("gn_fused", 1)
eng.set_option("gn_fused", 1)

for (N, Cc, H) in [(16, 320, 64), (16, 640, 32), (16, 1280, 16), (16, 1280, 8), (8, 1024, 16)]:
    row = {"N": N, "C": Cc, "H": H}
    for mode in (0, 1):
        eng.set_option("ln_vec", mode)
        ms = eng.bench_norm(1, N, Cc, H, H)
        gb = 2.0 * N * Cc * H * H * 2 / 1e9
        row["vec" if mode else "warp"] = ms * 1e3
        row["frac_vec" if mode else "frac_warp"] = gb / (ms * 1e-3) / HBM
    print(f"layernorm rows={N*H*H} C={Cc}: warp-per-row {row['warp']:7.1f} us ({row['frac_warp']:.2f} of HBM)   "
          f"vectorised {row['vec']:7.1f} us ({row['frac_vec']:.2f} of HBM)", flush=True)
    out["layernorm"].append(row)
eng.set_option("ln_vec", 1)

os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/micro_bench.json", "w"), indent=1)
print("MICRO DONE")
