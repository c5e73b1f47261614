"""Parity ON THE BENCHMARK CONFIGURATION (run with -m gpu on the B200 box): smb-vision-base width (768 / 12 heads, decoder
384 / 6 heads, MLP 3072 / 1536) at 512x512x320 = 20 480 tokens — the shapes bench.py times — against the CPU oracle
(oracle/videomae_oracle.py, fp32, `ATTN_IMPL="sdpa"`: the eager path would need a 20 GB score tensor per layer).

Depth is cut to 2 encoder + 1 decoder layers so that the oracle's forward + backward finishes in seconds on the box's
host cores; every kernel of the full model runs at its full-size shape (N = 20 480 / 7 168 tokens, d = 768 / 384, 5-D TMA
patch boxes over the whole volume, split last waves of the attention grid, 13 312-row loss), which is where a
tile-scheduling or tensor-map bug that the 216-token configurations cannot see would show.  Q / K weights are scaled up
(x3 each) so that the attention is peaked (score std ~ 2.7 instead of 0.3): with near-uniform attention an error in the
softmax would hide below the tolerance.

Tolerances: mask exact (sha256 known answer); loss rel <= 1e-4; logits Frobenius-rel <= 1e-2, max-abs-rel
<= 2e-2; embeddings <= 2e-2 / 5e-2; gradients Frobenius-rel <= 2e-2.
"""
import hashlib

import numpy as np
import pytest
import torch

import __graft_entry__ as ge
from oracle import videomae_oracle as vo
from oracle.mim_mask import OracleMaskGenerator

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def frob(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return (torch.linalg.norm(a - b) / torch.linalg.norm(b)).item()


def maxrel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).abs().max() / b.abs().max()).item()


@pytest.fixture(scope="module")
def ops():
    assert torch.cuda.is_available()
    from smb_vision_b200 import load, ops as _ops

    assert load().smbv_device_ok() == 0
    return _ops


@pytest.fixture(scope="module")
def base_cut():
    """smb-vision-base width at 512x512x320, 2 encoder + 1 decoder layers; peaked attention."""
    cfgd = dict(num_hidden_layers=2, decoder_num_hidden_layers=1)
    cfg = vo.OracleConfig(**cfgd)
    assert (cfg.hidden_size, cfg.num_attention_heads, cfg.decoder_hidden_size, cfg.num_patches) == (768, 12, 384, 20480)
    sd = vo.synthetic_state_dict(cfg, 1234)
    for k in sd:
        if k.endswith("query.weight") or k.endswith("key.weight"):
            sd[k] = sd[k] * 3.0
    x = vo.synthetic_volume(cfg, 1, 7)
    hc = ge.hf_config({k: getattr(cfg, k) for k in cfg.__dataclass_fields__})
    return cfg, sd, x, hc


@pytest.fixture(scope="module")
def sdpa_oracle():
    torch.set_num_threads(max(1, torch.get_num_threads()))
    old = vo.ATTN_IMPL
    vo.ATTN_IMPL = "sdpa"
    yield
    vo.ATTN_IMPL = old


def test_full_size_embeddings_match_oracle(ops, base_cut, sdpa_oracle):
    """`model.videomae(x).last_hidden_state` (reference :537-658) at N = 20 480, d = 768, 12 heads."""
    from smb_vision_b200.modeling import B200VideoMAEModel

    cfg, sd, x, hc = base_cut
    model = B200VideoMAEModel(hc).to(DEV)
    model.load_state_dict({k[len("videomae."):]: v for k, v in sd.items() if k.startswith("videomae.")}, strict=True)
    with torch.no_grad():
        emb = model(x.to(DEV)).last_hidden_state
    torch.cuda.synchronize()
    with torch.no_grad():
        ref = vo.encoder(sd, cfg, x, None)
    assert emb.shape == ref.shape == (1, 20480, 768) and emb.dtype == torch.float32
    f, m = frob(emb, ref), maxrel(emb, ref)
    print(f"full-size embeddings: frob-rel {f:.3e}, max-abs-rel {m:.3e}")
    assert f <= 2e-2 and m <= 5e-2, (f, m)
    # negative control: the comparison sees the attention — with uniform attention (Q = 0) the output moves far more than the tolerance
    sd0 = dict(sd)
    sd0["videomae.encoder.layer.0.attention.attention.query.weight"] = torch.zeros_like(sd["videomae.encoder.layer.0.attention.attention.query.weight"])
    sd0["videomae.encoder.layer.0.attention.attention.q_bias"] = torch.zeros_like(sd["videomae.encoder.layer.0.attention.attention.q_bias"])
    model.load_state_dict({k[len("videomae."):]: v for k, v in sd0.items() if k.startswith("videomae.")}, strict=True)
    with torch.no_grad():
        moved = frob(model(x.to(DEV)).last_hidden_state, ref)
    print(f"  negative control (layer-0 attention made uniform): frob-rel {moved:.3e}")
    assert moved > 4e-2


def test_full_size_mim_forward_backward_match_oracle(ops, base_cut, sdpa_oracle):
    """`model(x, mask)` loss / logits and `loss.backward()` gradients (reference :753-908) with the seed-0 full-size mask."""
    from smb_vision_b200.modeling import B200VideoMAEForPreTraining

    cfg, sd, x, hc = base_cut
    np.random.seed(0)
    fine = OracleMaskGenerator(512, 320, 32, 16, 0.65)()
    assert hashlib.sha256(fine.astype(np.uint8).tobytes()).hexdigest()[:16] == "4a598b65ed9ea8db"  # SURVEY.md §8c known answer
    mask = torch.from_numpy(fine)[None]
    assert int(mask.sum()) == 13312
    model = B200VideoMAEForPreTraining(hc).to(DEV)
    model.load_state_dict(sd, strict=True)
    out = model(x.to(DEV), mask)
    out.loss.backward()
    torch.cuda.synchronize()
    sdg = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    loss, logits, _ = vo.pretrain_forward(sdg, cfg, x, mask)
    loss.backward()
    rel = abs(out.loss.item() - loss.item()) / loss.item()
    fl, ml = frob(out.logits.float(), logits.detach()), maxrel(out.logits.float(), logits.detach())
    print(f"full-size MIM: loss {out.loss.item():.6f} vs {loss.item():.6f} (rel {rel:.2e}); logits frob-rel {fl:.3e}, max-abs-rel {ml:.3e}")
    assert out.logits.shape == (1, 13312, 4096)
    assert rel <= 1e-4 and fl <= 1e-2 and ml <= 2e-2, (rel, fl, ml)
    errs = {k: frob(p.grad, sdg[k].grad) for k, p in model.named_parameters()}
    worst = sorted(errs.items(), key=lambda kv: -kv[1])[:5]
    print("  gradients, worst five:", ", ".join(f"{k} {v:.2e}" for k, v in worst))
    selection = ["videomae.embeddings.patch_embeddings.projection.weight", "videomae.embeddings.patch_embeddings.projection.bias",
                 "mask_token", "videomae.encoder.layer.0.attention.attention.q_bias", "videomae.encoder.layer.1.attention.attention.v_bias",
                 "videomae.encoder.layer.0.attention.attention.key.weight", "videomae.encoder.layer.1.intermediate.dense.weight",
                 "encoder_to_decoder.weight", "decoder.decoder_layers.0.attention.attention.query.weight",
                 "decoder.decoder_layers.0.output.dense.weight", "decoder.norm.weight", "decoder.head.weight", "decoder.head.bias"]
    bad = {k: errs[k] for k in selection if not errs[k] <= 2e-2}
    assert not bad, bad
    assert max(errs.values()) <= 5e-2, worst  # every one of the 53 parameters


# ---------------------------------------------------------------------------------------------------------------------
# kernel level, at the benchmark's shapes (moved here from tools/gpu_check.py so that the driver's `pytest -m gpu` runs them)
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("M,N,K,epi", [(20480, 2304, 768, "qkv"), (20480, 768, 768, "resid"), (20480, 3072, 768, "gelu"),
                                       (20480, 768, 3072, "resid"), (13312, 4096, 384, "bf16"), (7168, 3072, 768, "gelu"),
                                       (20480, 1152, 384, "qkv"), (7168, 384, 768, "f32")])
def test_full_size_gemm_with_fused_epilogues(ops, M, N, K, epi):
    """every projection of the model at its full-size shape, epilogue fused, vs an fp32 matmul of the same bf16 operands."""
    torch.manual_seed(M + N + K)
    a = torch.randn(M, K, device=DEV).bfloat16()
    w = (torch.randn(N, K, device=DEV) * 0.05).bfloat16()
    bias = torch.randn(N, device=DEV)
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        ref = a.float() @ w.float().t() + bias
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev
    if epi == "qkv":
        H = N // 192
        out = ops.gemm(a, w, bias, ops.EPI_QKV_HEADS, heads=H, tokens=M)  # [3,1,H,M,64]
        got, want = out.float(), ref.view(1, M, 3, H, 64).permute(2, 0, 3, 1, 4)
        tol = 5e-3
    elif epi == "resid":
        res = torch.randn(M, N, device=DEV)
        want = res + ref
        ops.gemm(a, w, bias, ops.EPI_RESID_F32, residual=res)
        got, tol = res, 1e-3
    elif epi == "gelu":
        got, want, tol = ops.gemm(a, w, bias, ops.EPI_GELU_BF16).float(), torch.nn.functional.gelu(ref), 5e-3
    elif epi == "bf16":
        got, want, tol = ops.gemm(a, w, bias, ops.EPI_BF16).float(), ref, 5e-3
    else:
        got, want, tol = ops.gemm(a, w, bias, ops.EPI_F32), ref, 1e-3
    assert frob(got, want) <= tol


def test_full_size_patch_embed_matches_fp32_reference(ops):
    """implicit-GEMM patch embedding over the whole 512x512x320 volume (all 20 480 tokens, then the visible-row compaction
    of the seed-0 mask) vs patchify + fp32 matmul + bias + position table (reference :124-139, :179-192)."""
    cfg = vo.OracleConfig()
    x = vo.synthetic_volume(cfg, 1, 7)
    torch.manual_seed(3)
    w = torch.randn(768, 4096) * 0.02
    b = torch.randn(768) * 0.1
    P = vo.patchify(x, cfg)[0].to(DEV)
    pos = vo.sinusoid_table(cfg.num_patches, 768)[0].to(DEV)
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        ref = P @ w.to(DEV).t() + b.to(DEV) + pos
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev
    xg = x[:, :, 0].contiguous().to(DEV)
    out = ops.patch_embed_fwd(xg, w.to(DEV), b.to(DEV), pos)
    assert frob(out[0], ref) <= 5e-3  # bf16 operands vs the fp32 product
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        ref16 = P.bfloat16().float() @ w.to(DEV).bfloat16().float().t() + b.to(DEV) + pos  # the same operand rounding
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev
    assert frob(out[0], ref16) <= 2e-5
    np.random.seed(0)
    mask = torch.from_numpy(OracleMaskGenerator(512, 320, 32, 16, 0.65)())
    fine = mask.to(torch.uint8)[None].to(DEV)
    _, _, slot, _ = ops.mask_index(fine)
    outv = ops.patch_embed_fwd(xg, w.to(DEV), b.to(DEV), pos, fine, slot, 7168)
    assert frob(outv[0], ref16[~mask.to(DEV)]) <= 2e-5


@pytest.mark.parametrize("kind", ["mse", "l1"])
def test_full_size_normpix_loss_and_gradient(ops, kind):
    """13 312 masked patches x 4 096 voxels: loss rel <= 1e-5 and dlogits vs autograd over the oracle's labels."""
    cfg = vo.OracleConfig()
    x = vo.synthetic_volume(cfg, 1, 7)
    np.random.seed(0)
    mask = torch.from_numpy(OracleMaskGenerator(512, 320, 32, 16, 0.65)())[None]
    lab = vo.labels_normpix(x, cfg)[mask].reshape(1, 13312, -1)
    logits = (0.3 * torch.randn(1, 13312, 4096, generator=torch.Generator().manual_seed(5))).bfloat16()
    _, midx, _, _ = ops.mask_index(mask.to(torch.uint8).to(DEV))
    loss, dl = ops.normpix_loss(x[:, :, 0].contiguous().to(DEV), midx, 13312, logits.to(DEV), True, 0 if kind == "mse" else 1)
    lf = logits.float().requires_grad_(True)
    ref = torch.nn.functional.mse_loss(lf, lab) if kind == "mse" else torch.nn.functional.l1_loss(lf, lab)
    ref.backward()
    assert abs(loss.item() - ref.item()) / ref.item() <= 1e-5
    assert frob(dl.float(), lf.grad) <= 4e-3
