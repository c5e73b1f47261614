"""GPU: the example entry points that stand where src/run_mim.py and src/run_inference.py / run_inspect.py stand
(small model, synthetic int16 volumes): the loop trains (loss falls), checkpoints load into the upstream class, the
extractor writes reference-format parquet / npy files, resumes, and matches a direct model call."""
import os
import sys

import numpy as np
import pandas as pd
import pytest
import torch

import __graft_entry__ as ge

pytestmark = pytest.mark.gpu
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "examples"))
OVR = ",".join(f"{k}={ge.SMALL64[k]}" for k in ["hidden_size", "num_hidden_layers", "num_attention_heads", "intermediate_size", "decoder_hidden_size",
                                              "decoder_num_hidden_layers", "decoder_num_attention_heads", "decoder_intermediate_size"])


@pytest.mark.parametrize("graph", [False, True])
def test_train_mim_example_learns_and_checkpoints(tmp_path, graph):
    import transformers

    import train_mim

    out = str(tmp_path / "ckpt")
    losses = train_mim.main(["--synthetic", "2", "--image_size", "96", "--depth", "96", "--steps", "12", "--batch", "2", "--learning_rate", "1e-3",
                             "--warmup_ratio", "0.1", "--config_overrides", OVR, "--output_dir", out] + (["--cuda_graph"] if graph else []))
    assert len(losses) == 12 and all(np.isfinite(losses)) and losses[-1] < losses[0]
    up = transformers.VideoMAEForPreTraining.from_pretrained(out)  # the reference's class loads what we saved
    assert up.config.hidden_size == 128 and os.path.exists(os.path.join(out, "optimizer.pt"))


def test_extract_embeddings_example_formats_and_resume(tmp_path):
    import extract_embeddings
    from smb_vision_b200.data import VolumePreprocessor
    from smb_vision_b200.modeling import B200VideoMAEModel

    save = str(tmp_path / "emb")
    common = ["--image_size", "96", "--depth", "96", "--config_overrides", OVR, "--save_dir", save]
    files = extract_embeddings.main(["--synthetic", "5", "--format", "parquet", "--model_id", "small64"] + common)
    assert len(files) == 5
    df = pd.read_parquet(os.path.join(save, "model_id=small64", "synthetic_0003.parquet"))
    assert list(df["embedding_shape"][0]) == [216, 128] and df["uid"][0] == "synthetic_0003"
    # same numbers as a direct call on the same prepared volume
    g = torch.Generator().manual_seed(0)
    raws = [torch.randint(-1100, 1500, (96, 96, 96), generator=g, dtype=torch.int16) for _ in range(5)]
    hc = ge.hf_config(ge.SMALL64)
    torch.manual_seed(0)
    model = B200VideoMAEModel(hc).to("cuda:0").eval()
    x = VolumePreprocessor(96, 96, device="cuda:0")(raws[3]).unsqueeze(0)
    with torch.no_grad():
        want = model(x).last_hidden_state[0].cpu().numpy()
    assert np.array_equal(np.asarray(df["embedding"][0]).reshape(216, 128), want)
    assert extract_embeddings.main(["--synthetic", "5", "--format", "parquet", "--model_id", "small64"] + common) == []  # resume: nothing left
    npys = extract_embeddings.main(["--synthetic", "2", "--format", "npy", "--save_dir", str(tmp_path / "npy")] + common[:-2])
    assert [os.path.basename(p) for p in npys] == ["synthetic_0000.npy", "synthetic_0001.npy"]
    assert np.load(npys[0]).shape == (1, 216, 128)


def test_vjepa_step_with_fused_optimiser_and_ema_matches_torch_loop():
    """V-JEPA step (first slice): upstream VJEPA2Model through the attention plug-in + FusedAdamW on a flat
    arena + the one-launch EMA target update, against the reference's own loop in plain torch (sdpa attention, clip_grad_norm_,
    torch.optim.AdamW with Trainer's decay groups, per-parameter `mul_().add_()` EMA — src/run_vjepa.py:87-137)."""
    import copy

    import transformers
    from transformers.models.vjepa2.modeling_vjepa2 import apply_masks

    import smb_vision_b200.attention_interface as ai
    import train_vjepa
    from smb_vision_b200.optim import EmaTarget, FusedAdamW

    dev = torch.device("cuda", 0)
    name = ai.register()

    def build(attn):
        c = transformers.VJEPA2Config(patch_size=16, crop_size=64, frames_per_clip=64, tubelet_size=16, in_chans=1, hidden_size=128,
                                      num_attention_heads=2, num_hidden_layers=2, pred_hidden_size=64, pred_num_attention_heads=2,
                                      pred_num_hidden_layers=2, pred_num_mask_tokens=2)
        c._attn_implementation = attn
        torch.manual_seed(0)
        return transformers.VJEPA2Model(c).to(dev).train()

    g = torch.Generator().manual_seed(1)
    batches = []
    for _ in range(3):
        x = torch.rand(2, 64, 1, 64, 64, generator=g).to(dev)
        perm = torch.randperm(64, generator=g)
        batches.append((x, [perm[:40].sort().values[None].repeat(2, 1).to(dev)], [perm[40:].sort().values[None].repeat(2, 1).to(dev)]))

    # ---- reference loop (plain torch) ----
    ma = build("sdpa")
    ta = copy.deepcopy(ma)
    for p in ta.parameters():
        p.requires_grad = False
    nd = lambda n: ("bias" in n or "norm" in n)
    opt_a = torch.optim.AdamW([{"params": [p for n, p in ma.named_parameters() if not nd(n)], "weight_decay": 0.01},
                               {"params": [p for n, p in ma.named_parameters() if nd(n)], "weight_decay": 0.0}], lr=1e-3)
    losses_a = []
    for x, ctx, tgt in batches:
        opt_a.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            out = ma(pixel_values_videos=x, context_mask=ctx, target_mask=tgt)
            with torch.no_grad():
                th = apply_masks(ta(pixel_values_videos=x, context_mask=ctx, target_mask=tgt, skip_predictor=True).last_hidden_state, tgt)
        loss = torch.nn.functional.l1_loss(out.predictor_output.last_hidden_state.float(), th.float())
        loss.backward()
        torch.nn.utils.clip_grad_norm_(ma.parameters(), 1.0)
        opt_a.step()
        with torch.no_grad():
            for pq, pk in zip(ma.parameters(), ta.parameters()):
                pk.data.mul_(0.99925).add_(pq.data, alpha=1.0 - 0.99925)
        losses_a.append(loss.item())

    # ---- B200 path ----
    mb = build(name)
    opt_b = FusedAdamW(mb, lr=1e-3, weight_decay=0.01, max_grad_norm=1.0)
    grads = opt_b.grad_arena()
    tb = EmaTarget(mb, momentum=0.99925)
    losses_b = [float(train_vjepa.vjepa_step(mb, tb, opt_b, grads, *b)) for b in batches]
    for a, b in zip(losses_a, losses_b):
        assert abs(a - b) / a <= 2e-2, (losses_a, losses_b)
    assert losses_b[-1] != losses_b[0]

    def frob(u, v):
        return ((u.double() - v.double()).norm() / v.double().norm()).item()

    # weight matrices only: the zero-initialised biases move by +-lr per step (Adam's first steps are sign-like), so their
    # relative distance is dominated by the sign of near-zero gradients under bf16 noise
    # (the same goes for the zero-initialised predictor mask tokens)
    big = lambda t: t.dim() >= 2 and float(t.detach().abs().mean()) > 5e-3
    worst = max(frob(pb.detach(), pa.detach()) for pa, pb in zip(ma.parameters(), mb.parameters()) if big(pa))
    worst_t = max(frob(pb.detach(), pa.detach()) for pa, pb in zip(ta.parameters(), tb.model.parameters()) if big(pa))
    assert worst <= 5e-2 and worst_t <= 1e-4, (worst, worst_t)
    assert not any(p.requires_grad for p in tb.model.parameters())
    # the EMA launch alone is bit-exact with the per-parameter torch loop
    src = torch.randn(4096, device=dev)
    t1 = torch.randn(4096, device=dev)
    t2 = t1.clone()
    t1.mul_(0.99925).add_(src, alpha=1.0 - 0.99925)
    import ctypes as C
    from smb_vision_b200 import ops
    from smb_vision_b200._lib import call
    call("smbv_ema_update", ops._ptr(t2), ops._ptr(src), 4096, 0.99925, float(1.0 - 0.99925), C.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert torch.equal(t1, t2)
    # the native target-encoder forward reads the momentum weights in place: it equals the torch forward of the same module
    # (plug-in attention, bf16 autocast) and follows an EMA update that torch's version counters do not see
    x, ctx, tgt = batches[0]
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        want = tb.model(pixel_values_videos=x, context_mask=ctx, target_mask=tgt, skip_predictor=True).last_hidden_state.float()
    got = tb.encode(x)
    assert got.dtype == torch.float32 and frob(got, want) <= 2e-2, frob(got, want)
    with torch.no_grad():
        for p in mb.parameters():
            p.mul_(1.5)
    for _ in range(200):  # 200 momentum updates towards 1.5x weights: the target moves by ~2 %
        tb.update()
    moved = tb.encode(x)
    assert frob(moved, got) > 1e-3
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        want2 = tb.model(pixel_values_videos=x, context_mask=ctx, target_mask=tgt, skip_predictor=True).last_hidden_state.float()
    assert frob(moved, want2) <= 2e-2, frob(moved, want2)


@pytest.mark.parametrize("native_online", [False, True])
def test_train_vjepa_example_runs(native_online):
    """examples/train_vjepa.py end to end: VJEPAMaskGenerator + vjepa_collate_fn batches, plug-in or native online encoder,
    native momentum target, FusedAdamW + EMA: finite losses that move."""
    import train_vjepa

    torch.manual_seed(0)
    losses = train_vjepa.main(["--steps", "4", "--image_size", "96", "--depth", "96"] + (["--native_online"] if native_online else []))
    assert len(losses) == 4 and all(np.isfinite(l) and l > 0 for l in losses) and losses[-1] != losses[0]
