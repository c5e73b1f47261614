"""Timeline of the two softmax warps (tile A / tile B, quad 0) of one CTA of the forward kernel: clock64 at the hand-offs."""
import ctypes as C, os, sys
os.environ["SMBV_ATTN_FWD_STAGGER_NS"] = "-1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from smb_vision_b200 import _lib
H, N = 12, 20480
torch.manual_seed(0)
q, k, v = (torch.randn(1, H, N, 64, device="cuda").to(torch.bfloat16) for _ in range(3))
o = torch.empty(1, N, H * 64, device="cuda", dtype=torch.bfloat16)
for _ in range(3):
    _lib.call("smbv_flash_attn_fwd_ex", C.c_void_p(q.data_ptr()), C.c_void_p(k.data_ptr()), C.c_void_p(v.data_ptr()), 1, H, N, 0.125,
              C.c_void_p(o.data_ptr()), None, 0, None, 0, C.c_void_p(torch.cuda.current_stream().cuda_stream))
buf = np.zeros((12, 32), dtype=np.int64)
lib = _lib.load()
lib.smbv_debug_read_fwd_trace.argtypes = [C.c_void_p]
assert lib.smbv_debug_read_fwd_trace(buf.ctypes.data_as(C.c_void_p)) == 0
t0 = buf[0, 0]
names = ["sfull", "ld_done", "exp_done", "pv_passed", "st_done"]
print("event        " + " ".join(f"j={j + 8:<5d}" for j in range(8)))
for t in range(2):
    for e, n in enumerate(names):
        print(f"{'AB'[t]}_{n:10s} " + " ".join(f"{buf[e + 6 * t, j] - t0:7d}" for j in range(8)))
for t in range(2):
    d = buf[6 * t:6 * t + 5, :24].astype(np.int64)
    print(f"tile {'AB'[t]}: period {np.diff(d[0]).mean():.0f}  wait_sfull {np.mean(d[0, 1:] - d[4, :-1]):.0f}  ld {np.mean(d[1] - d[0]):.0f}  max+exp {np.mean(d[2] - d[1]):.0f}  "
          f"wait_pv {np.mean(d[3] - d[2]):.0f}  st {np.mean(d[4] - d[3]):.0f}")
print("offset B-A at sfull:", (buf[6, :12] - buf[0, :12]).tolist())
