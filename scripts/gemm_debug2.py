"""clock64 breakdown for the HBM/epilogue-bound GEMMs (K=320 class) of the UNet (option gemm_debug)."""
import ctypes as C, math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import torch.nn.functional as F
from reface_b200.runtime import Engine
eng = Engine(0, arena_bytes=12 << 30)


def run(M, K, N, res=False, geglu=False, **opts):
    base = dict(gemm_pair=1, gemm_kmerge=1, gemm_bn=0, gemm_stages=0, gemm_persistent=1, gemm_debug=1)
    base.update(opts)
    for k, v in base.items():
        eng.set_option(k, v)
    x = torch.randn(M, K, device="cuda").half().float()
    w = (torch.randn(N, K, device="cuda") / math.sqrt(K)).half().float()
    b = torch.randn(N, device="cuda")
    r = torch.randn(M, N // 2 if geglu else N, device="cuda").half().float() if res else None
    y = eng.op_linear(x, w, b, residual=r, geglu=geglu)
    ref = F.linear(x, w, b)
    if geglu:
        a, gate = ref.chunk(2, dim=-1)
        ref = a * F.gelu(gate)
    if res:
        ref = ref + r
    err = float((y - ref).abs().max())
    best = 1e9
    for _ in range(3):
        eng.set_option("profile", 1)
        eng.op_linear(x, w, b, residual=r, geglu=geglu)
        ms, fl, n = eng.profile_read()
        eng.set_option("profile", 0)
        best = min(best, ms)
    buf = (C.c_ulonglong * (148 * 8))()
    eng._ck(eng.lib.rfb_debug_read(eng.h, buf, 148 * 8))
    d = np.array(buf[:], dtype=np.float64).reshape(148, 8)
    lead, peer = d[0::2], d[1::2]
    tot = max(lead[:, 0].mean(), 1)
    et = max(lead[:, 5].mean(), 1)
    byts = (M * K + (M * N // (2 if geglu else 1)) * (2 if res else 1)) * 2
    print(f"M={M} K={K} N={N} res={int(res)} geglu={int(geglu)} {opts}: {best*1e3:7.1f} us {2.0*M*N*K/best/1e9:7.1f} TF/s "
          f"{byts/best/1e6:7.1f} GB/s err={err:.1e} | MMA wait_full {100*lead[:,1].mean()/tot:5.1f}% wait_acc {100*lead[:,2].mean()/tot:5.1f}% "
          f"{tot/max(lead[:,7].mean(),1):6.0f} cyc/tile | epi warp2: wait {100*lead[:,4].mean()/et:5.1f}% prefetch {100*peer[:,0].mean()/max(peer[:,5].mean(),1):5.1f}%",
          flush=True)


for args in [dict(M=65536, K=320, N=320), dict(M=65536, K=320, N=320, res=True), dict(M=65536, K=320, N=960),
             dict(M=65536, K=320, N=2560, geglu=True), dict(M=65536, K=1280, N=320, res=True),
             dict(M=65536, K=2880, N=320, res=True), dict(M=65536, K=2880, N=320),
             dict(M=16384, K=640, N=640, res=True), dict(M=16384, K=640, N=5120, geglu=True),
             dict(M=16384, K=5760, N=640, res=True), dict(M=4096, K=11520, N=1280, res=True)]:
    run(**args)
print("DEBUG2 DONE")
