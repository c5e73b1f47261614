"""V-JEPA2-3D ViT-L encoder forward (the embedding-extraction / momentum-target-encoder pass) at
512x512x320 = 20480 tokens, batch 1: the native encoder vs the upstream transformers model in bf16 with torch SDPA on the
same box.  usage: python tools/run_vjepa.py [steps] [upstream=1]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from transformers import VJEPA2Config, VJEPA2Model
from smb_vision_b200.vjepa import B200VJEPA2Model

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
with_upstream = (sys.argv[2] != "0") if len(sys.argv) > 2 else True
dev = torch.device("cuda", 0)
c = VJEPA2Config(patch_size=16, crop_size=512, frames_per_clip=320, tubelet_size=16, in_chans=1)  # src/run_vjepa.py:220-232 on ViT-L
torch.manual_seed(0)
with torch.device(dev):
    model = B200VJEPA2Model(c, with_predictor=False).eval()
x = torch.rand(1, 320, 1, 512, 512, device=dev)
N, d, L = 20480, c.hidden_size, c.num_hidden_layers
flops = 2 * N * 4096 * d + L * (2 * N * d * 12 * d + 4 * N * N * d)


def timed(fn, n):
    for _ in range(2):
        out = fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, out


ms, emb = timed(lambda: model.get_vision_features(x), steps)
print(f"native V-JEPA ViT-L encoder (N={N}): {ms:.2f} ms = {1e3 / ms:.2f} volumes/s, {flops / ms / 1e9:.0f} TFLOP/s, "
      f"finite={bool(torch.isfinite(emb).all())}, peak mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB")
if with_upstream:
    c2 = VJEPA2Config(patch_size=16, crop_size=512, frames_per_clip=320, tubelet_size=16, in_chans=1)
    c2._attn_implementation = "sdpa"
    with torch.device(dev):
        up = VJEPA2Model(c2).eval()
    sd = {k: v for k, v in model.state_dict().items()}
    sd = {k.replace("patch_embeddings.proj_3d", "patch_embeddings.proj"): v for k, v in sd.items()}
    print("upstream load:", up.load_state_dict(sd, strict=False).unexpected_keys)

    def ref():
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
            return up(pixel_values_videos=x, skip_predictor=True).last_hidden_state

    ms_u, emb_u = timed(ref, max(2, steps // 2))
    rel = (torch.linalg.norm(emb.double() - emb_u.double()) / torch.linalg.norm(emb_u.double())).item()
    print(f"upstream transformers VJEPA2Model, bf16 autocast + SDPA: {ms_u:.2f} ms = {1e3 / ms_u:.2f} volumes/s; "
          f"native vs upstream Frobenius-rel {rel:.3e}; speed-up {ms_u / ms:.2f}x")

# ---- online-encoder training pass: forward + backward of the encoder alone (loss = mean(last_hidden_state^2)) ----
if len(sys.argv) > 3 and sys.argv[3] == "train":
    model.train()

    def native_step():
        model.zero_grad(set_to_none=True)
        seq = model._runner.differentiable(x)
        seq.pow(2).mean().backward()
        return seq

    torch.cuda.reset_peak_memory_stats()
    ms_t, _ = timed(native_step, max(2, steps // 2))
    print(f"native V-JEPA ViT-L encoder forward + backward: {ms_t:.1f} ms = {3 * flops / ms_t / 1e9:.0f} TFLOP/s (3x forward flops), "
          f"peak mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB, grads finite="
          f"{all(bool(torch.isfinite(p.grad).all()) for p in model.encoder.parameters())}")
    if with_upstream:
        up.train()

        def up_step():
            up.zero_grad(set_to_none=True)
            with torch.autocast("cuda", dtype=torch.bfloat16):
                seq = up.encoder(x).last_hidden_state
            seq.float().pow(2).mean().backward()
            return seq

        ms_ut, _ = timed(up_step, max(2, steps // 2))
        print(f"upstream encoder forward + backward, bf16 autocast + SDPA: {ms_ut:.1f} ms; speed-up {ms_ut / ms_t:.2f}x")
