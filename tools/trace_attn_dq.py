"""Timeline of one CTA of the dQ kernel (SMBV_DQ_KNOCK=7): clock64 at the main events, relative, per key block."""
import ctypes as C, os, sys
os.environ["SMBV_DEV_HOOKS"] = "1"
os.environ["SMBV_DQ_KNOCK"] = "7"; os.environ["SMBV_SKIP_DKDV"] = "1"; os.environ["SMBV_ATTN_BWD_OVERLAP"] = "0"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from smb_vision_b200 import ops, _lib
H, N = 6, 20480
torch.manual_seed(0)
q, k, v = (torch.randn(H, N, 64, device="cuda").to(torch.bfloat16) for _ in range(3))
dout = torch.randn(N, H * 64, device="cuda").to(torch.bfloat16)
o, lse = ops.flash_attn_fwd(q[None], k[None], v[None], 0.125, return_lse=True)
for _ in range(3):
    ops.flash_attn_bwd(q, k, v, o[0], dout, lse[0], 0.125)
buf = np.zeros((16, 24), dtype=np.int64)
lib = _lib.load()
lib.smbv_debug_read_dq_trace.argtypes = [C.c_void_p]
assert lib.smbv_debug_read_dq_trace(buf.ctypes.data_as(C.c_void_p)) == 0
t0 = buf[1, 4]
names = ["tma_issued", "A_kv_ready", "A_sfree_h0", "A_sfree_h1", "B_pfull_h0", "B_pfull_h1", "M_sfull_h0", "M_sfull_h1", "M_ld_h0", "M_ld_h1",
         "M_math_h0", "M_math_h1", "M_pd_h0", "M_pd_h1", "M_st_h0", "M_st_h1"]
print("event          " + " ".join(f"j={j:<5d}" for j in range(4, 12)))
for e, n in enumerate(names):
    print(f"{n:14s} " + " ".join(f"{buf[e, j] - t0:7d}" for j in range(4, 12)))
print("period per block (A_sfree_h0):", np.diff(buf[2, 4:20]).tolist())
