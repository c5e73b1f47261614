"""Run the fc2-dgrad (dGELU epilogue) or the training fc1 (GELU + saved pre-activation) GEMM alone (for ncu).
usage: python tools/run_gemm_dgelu.py M d m dgelu|gelupre"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from smb_vision_b200 import ops
M, d, m = (int(v) for v in sys.argv[1:4]); kind = sys.argv[4] if len(sys.argv) > 4 else "dgelu"
x = torch.randn(M, d, device="cuda").bfloat16()
w1 = (torch.randn(m, d, device="cuda") * 0.05).bfloat16()
w2 = (torch.randn(d, m, device="cuda") * 0.05).bfloat16()
bm = torch.randn(m, device="cuda")
pre = torch.randn(M, m, device="cuda").bfloat16()
f = torch.empty(M, m, device="cuda", dtype=torch.bfloat16)
run = (lambda: ops.linear_dgrad(x, w2, aux=pre)) if kind == "dgelu" else (lambda: ops.gemm_ex(x, w1, M, m, d, ops.EPI_GELU_BF16, f, bias=bm, aux=pre))
for _ in range(3): run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); run(); e1.record(); torch.cuda.synchronize()
print(f"{kind} {M}x{m}x{d}: {e0.elapsed_time(e1) * 1e3:.1f} us")
