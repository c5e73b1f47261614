"""A/B of the attention-forward kernel variants (env SMBV_ATTN_FOLD, read once per process -> one subprocess per mode):
parity against fp32 eager attention at small shapes (incl. ragged N and large score jumps) and CUDA-event timing at the three
model shapes.  usage: python tools/attn_ab.py [modes...]   (default: 1 2 0)"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child():
    import torch
    from smb_vision_b200 import ops

    torch.manual_seed(0)
    dev = "cuda"
    for B, H, N, amp in [(1, 2, 216, 1.0), (2, 3, 640, 1.0), (1, 2, 3000, 1.0), (1, 2, 2048, 3.0), (1, 1, 1, 1.0), (1, 1, 130, 5.0)]:
        q = (torch.randn(B, H, N, 64, device=dev) * amp).bfloat16()
        k = (torch.randn(B, H, N, 64, device=dev) * amp).bfloat16()
        v = torch.randn(B, H, N, 64, device=dev).bfloat16()
        out, lse = ops.flash_attn_fwd(q, k, v, 0.125, return_lse=True)
        s = (q.float() @ k.float().transpose(-1, -2)) * 0.125
        ref = (torch.softmax(s, -1) @ v.float()).transpose(1, 2).reshape(B, N, H * 64)
        fr = ((out.float() - ref).norm() / ref.norm()).item()
        le = (lse - torch.logsumexp(s, -1)).abs().max().item()
        print(f"  parity B{B} H{H} N{N} amp{amp}: frob {fr:.2e} lse_err {le:.2e} {'OK' if fr < 1e-2 and le < 2e-2 else 'FAIL'}", flush=True)
    for H, N in [(12, 20480), (6, 20480), (12, 7168)]:
        q, k, v = (torch.randn(1, H, N, 64, device=dev).bfloat16() for _ in range(3))
        for _ in range(3):
            ops.flash_attn_fwd(q, k, v, 0.125)
        torch.cuda.synchronize()
        best = 1e9
        for rep in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                ops.flash_attn_fwd(q, k, v, 0.125)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / 10)
        print(f"  time H{H} N{N}: {best:.4f} ms = {4.0 * N * N * 64 * H / best / 1e9:.0f} TFLOP/s", flush=True)


if __name__ == "__main__":
    if os.environ.get("ATTN_AB_CHILD"):
        child()
    else:
        for mode in (sys.argv[1:] or ["1", "2", "0"]):
            print(f"SMBV_ATTN_FOLD={mode}", flush=True)
            env = dict(os.environ, ATTN_AB_CHILD="1", SMBV_ATTN_FOLD=mode)
            r = subprocess.run([sys.executable, __file__], env=env, capture_output=True, text=True, timeout=300)
            print(r.stdout + (r.stderr[-1500:] if r.returncode else ""), flush=True)
