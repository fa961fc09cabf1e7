"""clock64 timeline of one CTA of attn_flash3_kernel (option gemm_debug): per KV tile, when each role passed its waits."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from reface_b200.runtime import Engine
eng = Engine(0, arena_bytes=8 << 30)
N, L, heads, d = 16, int(os.environ.get("L", 4096)), 8, int(os.environ.get("D", 40))
eng.set_option("attn_flash", int(os.environ.get("MODE", 4)))
eng.set_option("attn_poly", int(os.environ.get("POLY", 0)))
eng.set_option("attn_pingpong", int(os.environ.get("PINGPONG", 1)))
qkv = torch.randn(N, L, 3 * heads * d, device="cuda").half().float()
eng.op_attention(qkv, heads)
eng.set_option("gemm_debug", 1)
eng.op_attention(qkv, heads)
buf = (C.c_ulonglong * (148 * 8))()
eng._ck(eng.lib.rfb_debug_read(eng.h, buf, 148 * 8))
eng.set_option("gemm_debug", 0)
v = list(buf)
t0 = min(x for x in v[:32 * 16] if x)
names = {0: "K_issue", 1: "V_issue", 10: "kf_ok", 2: "S_issue", 12: "vf_ok", 3: "PV_issue", 4: "sfull_ok", 5: "S_loaded",
         6: "max_xchg", 13: "turn_ok", 7: "exp_done", 8: "pfree_ok", 9: "pfull_arr"}
order = [0, 1, 10, 2, 12, 3, 4, 5, 6, 13, 7, 8, 9]
o1 = [4, 6, 13, 7, 9]
print("tile " + " ".join(f"{names[k]:>9s}" for k in order) + " | second query tile (warp 10): " + " ".join(f"{names[k]:>9s}" for k in o1))
for j in range(min(L // 128, 32)):
    print(f"{j:4d} " + " ".join(f"{(v[j*16+k]-t0) if v[j*16+k] else -1:9d}" for k in order) + " | " +
          " ".join(f"{(v[512+j*16+k]-t0) if v[512+j*16+k] else -1:9d}" for k in o1))
