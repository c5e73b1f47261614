"""Bring-up / A-B tool of the attention backward: fused one-pass kernel vs the two-kernel deterministic path, parity against fp32
autograd at small shapes and against each other at the model shapes, CUDA-event timings.  usage: python tools/attn_bwd_ab.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from smb_vision_b200 import ops

DEV = "cuda"


def frob(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm()).item()


def timeit(fn, iters=5, warmup=2):
    for _ in range(warmup):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


for B, H, N in [(1, 1, 128), (1, 1, 256), (2, 2, 216), (3, 1, 130), (1, 2, 1000), (4, 12, 1960)]:
    g = torch.Generator(device=DEV).manual_seed(N)
    q, k, v = (torch.randn(B, H, N, 64, device=DEV, generator=g).bfloat16() for _ in range(3))
    dout = torch.randn(B, N, H * 64, device=DEV, generator=g).bfloat16()
    out, lse = ops.flash_attn_fwd(q, k, v, 0.125, return_lse=True)
    dq, dk, dv = ops.flash_attn_bwd(q, k, v, out, dout, lse, 0.125, deterministic=False)
    torch.cuda.synchronize()
    qf, kf, vf = (t.float().requires_grad_(True) for t in (q, k, v))
    p = torch.softmax(qf @ kf.transpose(-1, -2) * 0.125, dim=-1)
    ref = (p @ vf).transpose(1, 2).reshape(B, N, H * 64)
    ref.backward(dout.float())
    d2 = ops.flash_attn_bwd(q, k, v, out, dout, lse, 0.125, deterministic=True)
    print(f"B{B} H{H} N{N}: fused vs fp32 dq {frob(dq, qf.grad):.2e} dk {frob(dk, kf.grad):.2e} dv {frob(dv, vf.grad):.2e} | "
          f"two-kernel dq {frob(d2[0], qf.grad):.2e} dk {frob(d2[1], kf.grad):.2e} dv {frob(d2[2], vf.grad):.2e}", flush=True)

for H, N in [(6, 20480), (12, 7168), (12, 20480)]:
    torch.manual_seed(0)
    q, k, v = (torch.randn(H, N, 64, device=DEV).to(torch.bfloat16) for _ in range(3))
    dout = torch.randn(N, H * 64, device=DEV).to(torch.bfloat16)
    o, lse = ops.flash_attn_fwd(q[None], k[None], v[None], 0.125, return_lse=True)
    a = ops.flash_attn_bwd(q, k, v, o[0], dout, lse[0], 0.125, deterministic=False)
    b = ops.flash_attn_bwd(q, k, v, o[0], dout, lse[0], 0.125, deterministic=True)
    a2 = ops.flash_attn_bwd(q, k, v, o[0], dout, lse[0], 0.125, deterministic=False)
    torch.cuda.synchronize()
    t_f = timeit(lambda: ops.flash_attn_bwd(q, k, v, o[0], dout, lse[0], 0.125, deterministic=False))
    t_d = timeit(lambda: ops.flash_attn_bwd(q, k, v, o[0], dout, lse[0], 0.125, deterministic=True))
    fl = 8.0 * N * N * 64 * H
    print(f"H{H} N{N}: fused {t_f:.3f} ms ({fl / t_f / 1e9:.0f} TF/s of 8N^2dH) two-kernel {t_d:.3f} ms ({fl / t_d / 1e9:.0f}) | "
          f"fused vs two-kernel dq {frob(a[0], b[0]):.2e} dk {frob(a[1], b[1]):.2e} dv {frob(a[2], b[2]):.2e} | "
          f"run-to-run dq {frob(a2[0], a[0]):.1e} dk equal {torch.equal(a2[1], a[1])} dv equal {torch.equal(a2[2], a[2])}", flush=True)
