"""Oracle restatement of the reference VideoMAE-3D MIM forward (test infrastructure only).

A plain-torch, CPU, functional restatement of
``/root/reference/src/models/videomae/modeling_videomae.py`` for the hot path
(SURVEY.md §8 a′).  Every function cites the reference lines it follows.  It is
pinned by ``tests/golden/tiny_mim.npz`` / ``tiny_embed.npz``, which were produced by
the reference itself (``oracle/make_golden.py``).  Works in float32 or float64 and
is differentiable through torch autograd, so gradients can be checked too.

Never imported by the product package.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np
import torch
import torch.nn.functional as F


@dataclass
class OracleConfig:
    """Subset of ``transformers.VideoMAEConfig`` the path reads (defaults = smb-vision-base,
    after ``src/run_mim.py:322-330``)."""

    image_size: int = 512
    num_frames: int = 320
    patch_size: int = 16
    tubelet_size: int = 16
    num_channels: int = 1
    hidden_size: int = 768
    num_hidden_layers: int = 12
    num_attention_heads: int = 12
    intermediate_size: int = 3072
    layer_norm_eps: float = 1e-12
    decoder_hidden_size: int = 384
    decoder_num_hidden_layers: int = 4
    decoder_num_attention_heads: int = 6
    decoder_intermediate_size: int = 1536
    norm_pix_loss: bool = True
    use_mean_pooling: bool = True
    qkv_bias: bool = True

    @property
    def grid(self):
        return (self.num_frames // self.tubelet_size, self.image_size // self.patch_size, self.image_size // self.patch_size)

    @property
    def num_patches(self):
        g = self.grid
        return g[0] * g[1] * g[2]

    @property
    def patch_dim(self):
        return self.tubelet_size * self.patch_size * self.patch_size * self.num_channels

    @classmethod
    def from_hf(cls, cfg) -> "OracleConfig":
        return cls(**{k: getattr(cfg, k) for k in cls.__dataclass_fields__})


TINY = dict(
    image_size=96, num_frames=96, patch_size=16, tubelet_size=16, num_channels=1,
    hidden_size=64, num_hidden_layers=2, num_attention_heads=4, intermediate_size=128,
    decoder_hidden_size=32, decoder_num_hidden_layers=1, decoder_num_attention_heads=2,
    decoder_intermediate_size=64,
)


def sinusoid_table(n_position: int, d_hid: int) -> torch.Tensor:
    """modeling_videomae.py:95-106 — float64 numpy, sin on even j, cos on odd j, cast to f32."""
    pos = np.arange(n_position, dtype=np.float64)[:, None]
    j = np.arange(d_hid)
    ang = pos / np.power(10000, 2 * (j // 2) / d_hid)[None, :]
    ang[:, 0::2] = np.sin(ang[:, 0::2])
    ang[:, 1::2] = np.cos(ang[:, 1::2])
    return torch.from_numpy(ang).float().unsqueeze(0)


def patchify(x: torch.Tensor, cfg: OracleConfig) -> torch.Tensor:
    """[B,T,C,H,W] -> [B,N,K]; token n=(tz,ty,tx) z-major, k=(dz,dy,dx[,c]).

    modeling_videomae.py:839-857 (view / permute(0,1,4,6,2,5,7,3) / view); for C=1 this is
    also the Conv3d im2col order of :172-192."""
    B, T, C, H, W = x.shape
    ts, ps = cfg.tubelet_size, cfg.patch_size
    x = x.view(B, T // ts, ts, C, H // ps, ps, W // ps, ps)
    x = x.permute(0, 1, 4, 6, 2, 5, 7, 3).contiguous()
    return x.view(B, (T // ts) * (H // ps) * (W // ps), ts * ps * ps * C)


ATTN_IMPL = "eager"  # "sdpa" = the ALL_ATTENTION_FUNCTIONS["sdpa"] backend real runs use (eager needs [H,N,N] fp32)


def _layer(x, sd, pre, heads, eps):
    """One pre-LN block: modeling_videomae.py:405-431 (+ :258-296 attention, :300-315, :359-388 MLP)."""
    d = x.shape[-1]
    h = F.layer_norm(x, (d,), sd[pre + "layernorm_before.weight"], sd[pre + "layernorm_before.bias"], eps)
    a = pre + "attention.attention."
    q = F.linear(h, sd[a + "query.weight"], sd.get(a + "q_bias"))  # :264
    k = F.linear(h, sd[a + "key.weight"], None)  # :261-262 zero k bias
    v = F.linear(h, sd[a + "value.weight"], sd.get(a + "v_bias"))  # :263
    B, N, _ = q.shape
    hd = d // heads
    q, k, v = (t.view(B, N, heads, hd).transpose(1, 2) for t in (q, k, v))  # :253-256
    if ATTN_IMPL == "sdpa":  # :270-289 with config._attn_implementation == "sdpa"
        o = F.scaled_dot_product_attention(q, k, v, scale=hd**-0.5).transpose(1, 2).reshape(B, N, d)
    else:
        s = torch.matmul(q, k.transpose(-1, -2)) * (hd**-0.5)  # :207, scaling :251
        p = torch.softmax(s, dim=-1)  # :210 (fp32 softmax)
        o = torch.matmul(p, v).transpose(1, 2).reshape(B, N, d)  # :219-221, :291-292
    o = F.linear(o, sd[pre + "attention.output.dense.weight"], sd[pre + "attention.output.dense.bias"])  # :311
    x = x + o  # :420
    h = F.layer_norm(x, (d,), sd[pre + "layernorm_after.weight"], sd[pre + "layernorm_after.bias"], eps)  # :423
    h = F.gelu(F.linear(h, sd[pre + "intermediate.dense.weight"], sd[pre + "intermediate.dense.bias"]))  # :368-370 exact erf
    h = F.linear(h, sd[pre + "output.dense.weight"], sd[pre + "output.dense.bias"])  # :382
    return x + h  # :385


def embed(sd, cfg: OracleConfig, x: torch.Tensor, mask: torch.Tensor | None, prefix="videomae."):
    """VideoMAEEmbeddings.forward, modeling_videomae.py:124-139 (+ patch embed :179-192)."""
    if x.shape[2] != cfg.num_channels:  # :181-184
        raise ValueError("Make sure that the channel dimension of the pixel values match with the one set in the configuration.")
    if x.shape[3] != cfg.image_size or x.shape[4] != cfg.image_size:  # :185-188
        raise ValueError(f"Input image size ({x.shape[3]}*{x.shape[4]}) doesn't match model ({cfg.image_size}*{cfg.image_size}).")
    w = sd[prefix + "embeddings.patch_embeddings.projection.weight"]
    b = sd[prefix + "embeddings.patch_embeddings.projection.bias"]
    P = patchify(x, cfg)
    # Conv3d weight [D,C,kz,ky,kx]; patch k order is (dz,dy,dx,c) -> move C last
    wk = w.permute(0, 2, 3, 4, 1).reshape(w.shape[0], -1)
    E = F.linear(P.to(wk.dtype), wk, b)
    E = E + sinusoid_table(cfg.num_patches, cfg.hidden_size).to(E.dtype)  # :129-131
    if mask is not None:  # :134-137
        B, _, C = E.shape
        E = E[~mask].reshape(B, -1, C)
    return E


def encoder(sd, cfg: OracleConfig, x: torch.Tensor, mask: torch.Tensor | None = None, prefix="videomae."):
    """VideoMAEModel.forward, modeling_videomae.py:537-658: embeddings, 12 layers, NO final LN when
    use_mean_pooling (:517-520) -> last_hidden_state (the embedding-extraction API)."""
    h = embed(sd, cfg, x, mask, prefix)
    for i in range(cfg.num_hidden_layers):
        h = _layer(h, sd, f"{prefix}encoder.layer.{i}.", cfg.num_attention_heads, cfg.layer_norm_eps)
    if not cfg.use_mean_pooling:
        h = F.layer_norm(h, (cfg.hidden_size,), sd[prefix + "layernorm.weight"], sd[prefix + "layernorm.bias"], cfg.layer_norm_eps)
    return h


def labels_normpix(x: torch.Tensor, cfg: OracleConfig) -> torch.Tensor:
    """modeling_videomae.py:822-867 for C != 3 (no un-normalise): per-patch (x-mean)/(sqrt(var_unbiased)+1e-6)."""
    P = patchify(x, cfg)
    if not cfg.norm_pix_loss:
        return P
    mu = P.mean(dim=-1, keepdim=True)
    var = P.var(dim=-1, unbiased=True, keepdim=True)
    return (P - mu) / (var.sqrt() + 1e-6)


def pretrain_forward(sd, cfg: OracleConfig, x: torch.Tensor, mask: torch.Tensor, loss_kind: str = "mse"):
    """VideoMAEForPreTraining.forward, modeling_videomae.py:753-908.  Returns (loss, logits, extras)."""
    if mask is None:  # :807-808
        raise ValueError("One must provided a boolean mask ")
    B = x.shape[0]
    enc = encoder(sd, cfg, x, mask)  # :791
    z = F.linear(enc, sd["encoder_to_decoder.weight"])  # :801-803, no bias
    dd = cfg.decoder_hidden_size
    pe = sinusoid_table(cfg.num_patches, dd).to(z.dtype).expand(B, -1, -1)  # :809-810
    pos_vis = pe[~mask].reshape(B, -1, dd)  # :811
    pos_msk = pe[mask].reshape(B, -1, dd)  # :812
    xfull = torch.cat([z + pos_vis, sd["mask_token"].to(z.dtype) + pos_msk], dim=1)  # :815 visible first
    h = xfull
    for j in range(cfg.decoder_num_hidden_layers):  # :695
        h = _layer(h, sd, f"decoder.decoder_layers.{j}.", cfg.decoder_num_attention_heads, cfg.layer_norm_eps)
    nm = pos_msk.shape[1]
    h = h[:, -nm:]  # :717-718
    h = F.layer_norm(h, (dd,), sd["decoder.norm.weight"], sd["decoder.norm.bias"], 1e-5)  # :676, :721
    logits = F.linear(h, sd["decoder.head.weight"], sd["decoder.head.bias"])  # :722
    with torch.no_grad():
        lab = labels_normpix(x.to(logits.dtype), cfg)
        lab = lab[mask].reshape(B, -1, lab.shape[-1])  # :893-894
    if loss_kind == "mse":
        loss = F.mse_loss(logits, lab)  # :896-897
    elif loss_kind == "l1":  # north-star variant; nn.L1Loss only appears at src/run_vjepa.py:108
        loss = F.l1_loss(logits, lab)
    else:
        raise ValueError(loss_kind)
    return loss, logits, {"labels": lab, "encoder": enc, "decoder_in": xfull}


def pretrain_forward_simmim(sd, cfg: OracleConfig, x: torch.Tensor, mask: torch.Tensor, loss_kind: str = "l1"):
    """The north star's SimMIM reading of the same model ("SimMIM mask-token blending ... masked-L1 reconstruction loss"; SURVEY.md §0
    fact 1): instead of dropping the masked tokens (MAE, modeling_videomae.py:134-137) their patch embeddings are REPLACED by a learned
    encoder-width mask token — `torch.where(bool_masked_pos.unsqueeze(-1), mask_token, embeddings)` followed by the position add, the
    blend of src/models/dinov2/modeling_dinov2.py:104-107, :113 — the encoder and the decoder run over all N tokens in their natural
    order, and the head / loss see the masked rows (reference :717-722, :893-897 unchanged).  Extra parameter:
    `videomae.embeddings.mask_token` [1,1,hidden]; the decoder-width `mask_token` of the MAE path is unused.
    Returns (loss, logits, extras)."""
    if mask is None:
        raise ValueError("One must provided a boolean mask ")
    B = x.shape[0]
    w = sd["videomae.embeddings.patch_embeddings.projection.weight"]
    b = sd["videomae.embeddings.patch_embeddings.projection.bias"]
    P = patchify(x, cfg)
    E = F.linear(P.to(w.dtype), w.permute(0, 2, 3, 4, 1).reshape(w.shape[0], -1), b)
    E = torch.where(mask.unsqueeze(-1), sd["videomae.embeddings.mask_token"].to(E.dtype), E)  # dinov2 :104-107
    h = E + sinusoid_table(cfg.num_patches, cfg.hidden_size).to(E.dtype)  # dinov2 :113 / videomae :129-131
    for i in range(cfg.num_hidden_layers):
        h = _layer(h, sd, f"videomae.encoder.layer.{i}.", cfg.num_attention_heads, cfg.layer_norm_eps)
    if not cfg.use_mean_pooling:
        h = F.layer_norm(h, (cfg.hidden_size,), sd["videomae.layernorm.weight"], sd["videomae.layernorm.bias"], cfg.layer_norm_eps)
    dd = cfg.decoder_hidden_size
    h = F.linear(h, sd["encoder_to_decoder.weight"]) + sinusoid_table(cfg.num_patches, dd).to(h.dtype)  # natural token order
    for j in range(cfg.decoder_num_hidden_layers):
        h = _layer(h, sd, f"decoder.decoder_layers.{j}.", cfg.decoder_num_attention_heads, cfg.layer_norm_eps)
    h = h[mask].reshape(B, -1, dd)  # masked rows, ascending n
    h = F.layer_norm(h, (dd,), sd["decoder.norm.weight"], sd["decoder.norm.bias"], 1e-5)
    logits = F.linear(h, sd["decoder.head.weight"], sd["decoder.head.bias"])
    with torch.no_grad():
        lab = labels_normpix(x.to(logits.dtype), cfg)
        lab = lab[mask].reshape(B, -1, lab.shape[-1])
    loss = F.l1_loss(logits, lab) if loss_kind == "l1" else F.mse_loss(logits, lab)
    return loss, logits, {"labels": lab}


def classify_forward(sd, cfg: OracleConfig, x: torch.Tensor, additional_features=None, labels=None, num_labels: int = 2,
                     problem_type: str | None = None):
    """VideoMAEForVideoClassification.forward, modeling_videomae.py:943-1023 (the reference's variant with
    `additional_features`, :927-937, :979-987): encoder -> mean over tokens -> fc_norm -> [cat features] -> classifier
    -> MSE / CE / BCE (:995-1012).  Returns (loss or None, logits)."""
    h = encoder(sd, cfg, x, None)
    if cfg.use_mean_pooling:  # :974-975
        h = F.layer_norm(h.mean(1), (cfg.hidden_size,), sd["fc_norm.weight"], sd["fc_norm.bias"], 1e-5)
    else:
        h = h[:, 0]
    if additional_features is not None:  # :979-987
        if additional_features.shape[-1] != sd["classifier.weight"].shape[1] - cfg.hidden_size:
            raise ValueError(f"Expected additional_features of size {sd['classifier.weight'].shape[1] - cfg.hidden_size}, got {additional_features.shape[-1]}")
        h = torch.cat([h, additional_features.to(h.dtype)], dim=-1)
    logits = F.linear(h, sd["classifier.weight"], sd["classifier.bias"])  # :989
    loss = None
    if labels is not None:  # :992-1012
        if problem_type is None:
            problem_type = ("regression" if num_labels == 1 else
                            "single_label_classification" if labels.dtype in (torch.long, torch.int) else "multi_label_classification")
        if problem_type == "regression":
            loss = F.mse_loss(logits.squeeze(), labels.squeeze()) if num_labels == 1 else F.mse_loss(logits, labels)
        elif problem_type == "single_label_classification":
            loss = F.cross_entropy(logits.view(-1, num_labels), labels.view(-1))
        else:
            loss = F.binary_cross_entropy_with_logits(logits, labels)
    return loss, logits


def synthetic_cls_state_dict(cfg: OracleConfig, num_labels: int, n_features: int, seed: int = 1234) -> dict:
    """encoder weights of `synthetic_state_dict` + fc_norm + classifier (keys of VideoMAEForVideoClassification)."""
    sd = {k: v for k, v in synthetic_state_dict(cfg, seed).items() if k.startswith("videomae.")}
    g = torch.Generator().manual_seed(seed + 1)
    d = cfg.hidden_size
    sd["fc_norm.weight"] = 1.0 + 0.1 * torch.randn(d, generator=g)
    sd["fc_norm.bias"] = 0.02 * torch.randn(d, generator=g)
    sd["classifier.weight"] = 0.05 * torch.randn(num_labels, d + n_features, generator=g)
    sd["classifier.bias"] = 0.02 * torch.randn(num_labels, generator=g)
    return sd


# --------------------------------------------------------------------------------------
# deterministic synthetic inputs shared by the golden generator, the tests and bench.py
# --------------------------------------------------------------------------------------
def param_shapes(cfg: OracleConfig) -> dict:
    """Checkpoint ABI: key -> shape, in a fixed order."""
    d, m, dd, dm = cfg.hidden_size, cfg.intermediate_size, cfg.decoder_hidden_size, cfg.decoder_intermediate_size
    K = cfg.patch_dim
    s = {"mask_token": (1, 1, dd)}
    s["videomae.embeddings.patch_embeddings.projection.weight"] = (d, cfg.num_channels, cfg.tubelet_size, cfg.patch_size, cfg.patch_size)
    s["videomae.embeddings.patch_embeddings.projection.bias"] = (d,)

    def layer(pre, d, m):
        a = pre + "attention.attention."
        if cfg.qkv_bias:  # reference :242-251
            s[a + "q_bias"] = (d,)
            s[a + "v_bias"] = (d,)
        s[a + "query.weight"] = (d, d)
        s[a + "key.weight"] = (d, d)
        s[a + "value.weight"] = (d, d)
        s[pre + "attention.output.dense.weight"] = (d, d)
        s[pre + "attention.output.dense.bias"] = (d,)
        s[pre + "intermediate.dense.weight"] = (m, d)
        s[pre + "intermediate.dense.bias"] = (m,)
        s[pre + "output.dense.weight"] = (d, m)
        s[pre + "output.dense.bias"] = (d,)
        for ln in ("layernorm_before", "layernorm_after"):
            s[pre + ln + ".weight"] = (d,)
            s[pre + ln + ".bias"] = (d,)

    for i in range(cfg.num_hidden_layers):
        layer(f"videomae.encoder.layer.{i}.", d, m)
    if not cfg.use_mean_pooling:  # final encoder LayerNorm, reference :517-520
        s["videomae.layernorm.weight"] = (d,)
        s["videomae.layernorm.bias"] = (d,)
    s["encoder_to_decoder.weight"] = (dd, d)
    for j in range(cfg.decoder_num_hidden_layers):
        layer(f"decoder.decoder_layers.{j}.", dd, dm)
    s["decoder.norm.weight"] = (dd,)
    s["decoder.norm.bias"] = (dd,)
    s["decoder.head.weight"] = (K, dd)
    s["decoder.head.bias"] = (K,)
    return s


def synthetic_state_dict(cfg: OracleConfig, seed: int = 1234, perturb: bool = True) -> dict:
    """Seeded weights with the reference's init statistics (modeling_videomae.py:495-505: N(0,0.02)
    weights, LN (1,0)) but with the zero-initialised parameters (biases, q_bias/v_bias, mask_token)
    perturbed when ``perturb`` so that errors in how they are applied are visible."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for k, shp in param_shapes(cfg).items():
        if "layernorm" in k or k.startswith("decoder.norm"):
            base = 1.0 if k.endswith("weight") else 0.0
            t = torch.full(shp, base) + (0.1 * torch.randn(shp, generator=g) if perturb else 0.0)
        elif k.endswith("bias") or k == "mask_token":
            t = 0.02 * torch.randn(shp, generator=g) if perturb else torch.zeros(shp)
        else:
            t = 0.02 * torch.randn(shp, generator=g)
        sd[k] = t.float()
    return sd


def synthetic_volume(cfg: OracleConfig, batch: int = 1, seed: int = 7) -> torch.Tensor:
    """uniform [0,1) f32 [B,T,1,H,W] (post-ScaleIntensityRanged range, src/dataloader/mim.py:154-161)
    plus a smooth per-axis sinusoid so patches are not i.i.d. noise (scripts/preprocess/create_dummy_data.py:50-54)."""
    g = torch.Generator().manual_seed(seed)
    T, H = cfg.num_frames, cfg.image_size
    x = torch.rand(batch, T, cfg.num_channels, H, H, generator=g)
    z = torch.sin(torch.arange(T) * (2 * math.pi / 64)).view(1, T, 1, 1, 1)
    y = torch.cos(torch.arange(H) * (2 * math.pi / 48)).view(1, 1, 1, H, 1)
    w = torch.sin(torch.arange(H) * (2 * math.pi / 40)).view(1, 1, 1, 1, H)
    return (0.6 * x + 0.2 + 0.06 * (z + y + w)).clamp_(0.0, 1.0).float()
