"""V-JEPA2-3D encoder on the sm_100a kernels: the path the reference runs for embedding extraction
and for its momentum TARGET encoder — ``VJEPA2Model.forward(pixel_values_videos, context_mask, target_mask,
skip_predictor=True)`` (reference src/models/vjepa/modeling_vjepa.py:1071-1149, called under ``torch.no_grad()`` at
src/run_vjepa.py:128-135) and ``get_vision_features`` (:1151-1153).

Same constructor (a ``VJEPA2Config`` as ``src/run_vjepa.py:220-232`` builds it: ``in_chans=1``, ``tubelet_size = patch_size
= 16``), same parameter names and shapes (``encoder.embeddings.patch_embeddings.proj_3d``, ``encoder.layer.N.{norm1,
attention.{query,key,value,proj}, norm2, mlp.{fc1,fc2}}``, ``encoder.layernorm``, ``predictor.*``), same output class.

Encoder forward = tubelet embedding (implicit-GEMM kernel, no position table) -> L x [LayerNorm -> fused QKV GEMM with all
three biases, head-major -> ``smbv_rope3d`` in place on Q and K -> tcgen05 flash attention -> proj + residual -> LayerNorm
-> fc1 + GELU -> fc2 + residual] -> final LayerNorm.  The predictor (12 x 384/12, head_dim 32; reference :559-746) runs on the
same kernels (`VJepaPredictorRunner`: context gather, Linear, mask token, position sort as ONE index kernel, the blocks with the
sorted positions as rotary ids and head_dim 32 zero-padded onto the head_dim-64 attention kernels, LayerNorm, projection).
With gradients enabled the encoder and the predictor are ONE autograd node each with hand-written backward passes
(`VJepaEncoderRunner.backward`, `VJepaPredictorRunner.backward`: the VideoMAE block backward of `training.py` + the transposed
rotary map + K-bias gradient + the adjoint gathers), so the online model of `examples/train_vjepa.py --native_online` trains
entirely on the kernels.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional

import torch
from torch import nn

from . import ops
from ._lib import SmbvError
from .modeling import _PackedLayer, _PretrainedIO, _block_forward, _f32, _params_signature

try:
    from transformers.utils import ModelOutput as _OutputBase
except Exception:  # pragma: no cover
    _OutputBase = object


@dataclass
class VJEPA2WithMaskedInputPredictorOutput(_OutputBase):
    """reference modeling_vjepa.py:36-62 (the upstream class of transformers 5.x dropped `target_hidden_state`)."""

    last_hidden_state: torch.Tensor = None
    masked_hidden_state: Optional[torch.Tensor] = None
    hidden_states: Optional[tuple] = None
    attentions: Optional[tuple] = None
    target_hidden_state: Optional[torch.Tensor] = None


@dataclass
class VJEPA2WithMaskedInputModelOutput(_OutputBase):
    """reference modeling_vjepa.py:65-101."""

    last_hidden_state: torch.Tensor = None
    masked_hidden_state: Optional[torch.Tensor] = None
    target_hidden_state: Optional[torch.Tensor] = None
    hidden_states: Optional[tuple] = None
    attentions: Optional[tuple] = None
    predictor_output: Optional[VJEPA2WithMaskedInputPredictorOutput] = None


# ---- parameter containers (names == the reference checkpoint ABI, modeling_vjepa.py:105-125, :231-261, :412-452) ----
class _RopeAttention(nn.Module):
    def __init__(self, d, qkv_bias):
        super().__init__()
        self.query = nn.Linear(d, d, bias=qkv_bias)
        self.key = nn.Linear(d, d, bias=qkv_bias)
        self.value = nn.Linear(d, d, bias=qkv_bias)
        self.proj = nn.Linear(d, d)


class _MLP(nn.Module):
    def __init__(self, d, m):
        super().__init__()
        self.fc1 = nn.Linear(d, m)
        self.fc2 = nn.Linear(m, d)


class _VJepaLayer(nn.Module):
    def __init__(self, d, m, eps, qkv_bias):
        super().__init__()
        self.norm1 = nn.LayerNorm(d, eps=eps)
        self.attention = _RopeAttention(d, qkv_bias)
        self.norm2 = nn.LayerNorm(d, eps=eps)
        self.mlp = _MLP(d, m)


class _PatchEmbeddings3D(nn.Module):
    def __init__(self, config):
        super().__init__()
        t, p = config.tubelet_size, config.patch_size
        self.proj_3d = nn.Conv3d(config.in_chans, config.hidden_size, kernel_size=(t, p, p), stride=(t, p, p))


class _VJepaEmbeddings(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.patch_embeddings = _PatchEmbeddings3D(config)


class _VJepaEncoder(nn.Module):
    def __init__(self, config):
        super().__init__()
        d = config.hidden_size
        self.embeddings = _VJepaEmbeddings(config)
        self.layer = nn.ModuleList([_VJepaLayer(d, int(d * config.mlp_ratio), config.layer_norm_eps, config.qkv_bias)
                                    for _ in range(config.num_hidden_layers)])
        self.layernorm = nn.LayerNorm(d, eps=config.layer_norm_eps)


class _VJepaPredictorEmbeddings(nn.Module):  # reference modeling_vjepa.py:559-575
    def __init__(self, config):
        super().__init__()
        self.predictor_embeddings = nn.Linear(config.hidden_size, config.pred_hidden_size)
        self.mask_tokens = nn.Parameter(torch.zeros(config.pred_num_mask_tokens, 1, 1, config.pred_hidden_size))


class _VJepaPredictor(nn.Module):  # reference modeling_vjepa.py:629-657: parameter containers only, the math is VJepaPredictorRunner
    def __init__(self, config):
        super().__init__()
        pd = config.pred_hidden_size
        self.embeddings = _VJepaPredictorEmbeddings(config)
        self.layer = nn.ModuleList([_VJepaLayer(pd, int(pd * config.pred_mlp_ratio), config.layer_norm_eps, config.qkv_bias)
                                    for _ in range(config.pred_num_hidden_layers)])
        self.layernorm = nn.LayerNorm(pd, eps=config.layer_norm_eps)
        self.proj = nn.Linear(pd, config.hidden_size)


def _init_weights(module, std):
    """reference modeling_vjepa.py:1017-1041: trunc-normal(std) matrices, zero biases, LayerNorm (1, 0)."""
    for m in module.modules():
        if isinstance(m, (nn.Linear, nn.Conv3d)):
            nn.init.trunc_normal_(m.weight, mean=0.0, std=std)
            if m.bias is not None:
                nn.init.zeros_(m.bias)
        elif isinstance(m, nn.LayerNorm):
            nn.init.ones_(m.weight)
            nn.init.zeros_(m.bias)


def _pack_layer(layer: nn.Module, heads: int, eps: float) -> _PackedLayer:
    """fp32 masters -> bf16 operands; Q, K, V fused into one [3d, d] weight with bias [q; k; v] (K has a real bias here,
    unlike VideoMAE's zero K bias)."""
    a = layer.attention
    d = a.query.weight.shape[0]
    zeros = torch.zeros(d, dtype=torch.float32, device=a.query.weight.device)
    p = _PackedLayer()
    p.wqkv = ops.cast_bf16(torch.cat([_f32(a.query.weight), _f32(a.key.weight), _f32(a.value.weight)], 0))
    p.bqkv = torch.cat([_f32(l.bias) if l.bias is not None else zeros for l in (a.query, a.key, a.value)]).contiguous()
    p.wo, p.bo = ops.cast_bf16(_f32(a.proj.weight)), _f32(a.proj.bias)
    p.w1, p.b1 = ops.cast_bf16(_f32(layer.mlp.fc1.weight)), _f32(layer.mlp.fc1.bias)
    p.w2, p.b2 = ops.cast_bf16(_f32(layer.mlp.fc2.weight)), _f32(layer.mlp.fc2.bias)
    p.g1, p.be1 = _f32(layer.norm1.weight), _f32(layer.norm1.bias)
    p.g2, p.be2 = _f32(layer.norm2.weight), _f32(layer.norm2.bias)
    p.heads, p.eps, p.hd = heads, eps, d // heads
    return p


class _ScratchGrads:
    """What `training.block_backward` needs from a gradient arena, for parameters that live in a module we do not lay
    out ourselves: zero-initialised fp32 buffers by parameter name (the weight-gradient GEMMs accumulate), one fused
    [3d, d] / [3d] buffer per block for Q, K, V that `per_parameter` splits into the three parameters' gradients."""

    def __init__(self, named: Dict[str, nn.Parameter], device):
        self.named, self.device = named, device
        self.offsets = named  # block_backward only asks `name in arena.offsets`
        self.buf: Dict[str, torch.Tensor] = {}
        self.fused: Dict[str, torch.Tensor] = {}

    def g(self, name: str) -> torch.Tensor:
        if name not in self.buf:
            self.buf[name] = torch.zeros(self.named[name].shape, dtype=torch.float32, device=self.device)
        return self.buf[name]

    def _fused(self, key: str, shape) -> torch.Tensor:
        if key not in self.fused:
            self.fused[key] = torch.zeros(shape, dtype=torch.float32, device=self.device)
        return self.fused[key]

    def fused_qkv(self, prefix: str) -> torch.Tensor:
        d = self.named[prefix + "attention.query.weight"].shape[0]
        return self._fused(prefix + "w", (3 * d, d))

    def fused_qkv_bias(self, prefix: str) -> torch.Tensor:
        return self._fused(prefix + "b", (3 * self.named[prefix + "attention.query.weight"].shape[0],))

    def per_parameter(self, d: int) -> Dict[str, torch.Tensor]:
        out = dict(self.buf)
        for key, t in self.fused.items():
            prefix, kind = key[:-1], ("weight" if key.endswith("w") else "bias")
            for j, lin in enumerate(("query", "key", "value")):
                out[f"{prefix}attention.{lin}.{kind}"] = t[j * d:(j + 1) * d]
        return out


class _EncoderFunction(torch.autograd.Function):
    """last_hidden_state = encoder(pixel_values_videos) with the hand-written backward; the parameters are inputs so that
    autograd routes their gradients (into `.grad`, i.e. into the flat arena when `FusedAdamW.grad_arena()` assigned it)."""

    @staticmethod
    def forward(ctx, runner, pixel_values_videos, names, *params):
        vol = runner.volume(pixel_values_videos)
        seq, saved = runner.encode_train(vol)
        ctx.runner, ctx.vol, ctx.acts, ctx.names = runner, vol, saved, names
        return seq

    @staticmethod
    def backward(ctx, dseq):
        grads = ctx.runner.backward(ctx.vol, ctx.acts, dseq)
        ctx.acts = None
        return (None, None, None) + tuple(grads.get(n) for n in ctx.names)


class VJepaEncoderRunner:
    """Encoder forward on the kernels for ANY module laid out like the reference's ``VJEPA2Encoder`` (modeling_vjepa.py:
    488-546: ``embeddings.patch_embeddings.proj_3d`` — ``proj`` upstream —, ``layer[i].{norm1, attention.{query,key,value,
    proj}, norm2, mlp.{fc1,fc2}}``, ``layernorm``): our own containers, the reference's model, or the deep copy the trainer
    keeps as momentum target (``optim.EmaTarget.encode``).  Weights are packed (bf16 operands, fused QKV) once per parameter
    version; `invalidate()` forces a repack after an update that bypasses torch's version counters (our EMA kernel)."""

    def __init__(self, encoder: nn.Module, config):
        self.encoder, self.config = encoder, config
        self._packed, self._sig, self.version = None, None, 0

    def invalidate(self) -> None:
        self.version += 1

    def __deepcopy__(self, memo):
        # copy.deepcopy(model) (the trainer's momentum target, src/run_vjepa.py:104) must not duplicate the packed bf16 cache
        import copy

        return VJepaEncoderRunner(copy.deepcopy(self.encoder, memo), copy.deepcopy(self.config, memo))

    @property
    def grid_size(self) -> int:
        return self.config.crop_size // self.config.patch_size

    @property
    def grid_depth(self) -> int:
        return self.config.frames_per_clip // self.config.tubelet_size

    def check_config(self):
        c = self.config
        if c.patch_size != 16 or c.tubelet_size != 16:
            raise SmbvError("smb_vision_b200 implements patch_size = tubelet_size = 16 (src/run_vjepa.py:226-229 sets both)")
        if c.in_chans != 1:
            raise SmbvError("smb_vision_b200 implements single-channel CT/MR volumes (in_chans=1, src/run_vjepa.py:227)")
        if c.hidden_size // c.num_attention_heads != 64:
            raise SmbvError("the native V-JEPA encoder implements head_dim 64 (ViT-L 1024/16, ViT-H 1280/20, ViT-g 1408/22)")
        if getattr(c, "hidden_act", "gelu") != "gelu":
            raise SmbvError("only hidden_act='gelu' (exact erf) is implemented")

    def packed(self):
        sig = (_params_signature(self.encoder), self.version)
        if self._packed is None or sig != self._sig:
            c = self.config
            pe = self.encoder.embeddings.patch_embeddings
            proj = pe.proj_3d if hasattr(pe, "proj_3d") else pe.proj
            self._packed = dict(
                wpe=ops.cast_bf16(_f32(proj.weight).reshape(c.hidden_size, -1).contiguous()), bpe=_f32(proj.bias),  # bf16 operand
                layers=[_pack_layer(l, c.num_attention_heads, c.layer_norm_eps) for l in self.encoder.layer],
                g=_f32(self.encoder.layernorm.weight), b=_f32(self.encoder.layernorm.bias))
            self._sig = sig
        return self._packed

    def volume(self, pixel_values_videos: torch.Tensor) -> torch.Tensor:
        if pixel_values_videos is None:  # reference :1103-1104
            raise ValueError("You have to specify pixel_values_videos")
        if pixel_values_videos.dim() != 5:
            raise ValueError("pixel_values_videos must be [batch, frames, channels, height, width]")
        B, T, C, H, W = pixel_values_videos.shape
        if C != self.config.in_chans:
            raise ValueError(f"expected {self.config.in_chans} input channel(s), got {C}")
        dev = self.encoder.layernorm.weight.device
        return pixel_values_videos.to(device=dev, dtype=torch.float32, non_blocking=True).reshape(B, T, H, W).contiguous()

    def encode(self, vol: torch.Tensor) -> torch.Tensor:
        """fp32 volume [B,T,H,W] -> fp32 last_hidden_state [B, N, d] (reference VJEPA2Encoder.forward, :509-546).  Like the
        reference the token grid follows the INPUT size (ids = arange(N), row length = config grid_size, :297-316)."""
        self.check_config()
        pk = self.packed()
        X = ops.patch_embed_fwd(vol, pk["wpe"], pk["bpe"], None)
        rope = (self.grid_size, None, min(max(self.grid_size, self.grid_depth, vol.shape[1] // 16), 256))
        for p in pk["layers"]:
            _block_forward(X, p, rope)
        return ops.layernorm_fwd(X, pk["g"], pk["b"], self.config.layer_norm_eps).float()

    # ---- training (the ONLINE encoder: forward keeping activations, hand-written backward) ----
    def encode_train(self, vol: torch.Tensor):
        """`encode` out of place, keeping what `backward` needs.  Returns (last_hidden_state fp32, saved)."""
        from .training import block_forward_train

        self.check_config()
        # a differentiable forward never trusts torch's version counters: fused / foreach optimisers (HF Trainer's default)
        # update parameters without bumping them, and the packed bf16 operands would go stale
        self.invalidate()
        pk = self.packed()
        X = ops.patch_embed_fwd(vol, pk["wpe"], pk["bpe"], None)
        rope = (self.grid_size, None, min(max(self.grid_size, self.grid_depth, vol.shape[1] // 16), 256))
        blocks = []
        for p in pk["layers"]:
            X, sv = block_forward_train(X, p, rope)
            blocks.append(sv)
        y, mean, rstd = ops.layernorm_fwd(X, pk["g"], pk["b"], self.config.layer_norm_eps, save_stats=True)
        return y.float(), (pk, blocks, X, mean, rstd, rope)

    def backward(self, vol: torch.Tensor, saved, dseq: torch.Tensor) -> Dict[str, torch.Tensor]:
        """Gradient of every encoder parameter (names relative to the encoder module) given d(last_hidden_state):
        final LayerNorm -> blocks in reverse (attention backward kernels, transposed rotary map, fused QKV wgrad / dgrad)
        -> tubelet-embedding weight gradient over the im2col rows of all tokens."""
        from .training import SideQueue, block_backward, vjepa_block_names

        pk, blocks, X, mean, rstd, rope = saved
        named = dict(self.encoder.named_parameters())
        sc = _ScratchGrads(named, vol.device)
        sq = SideQueue(vol.device)  # weight / bias gradients on the second stream (training.SideQueue)
        d = self.config.hidden_size
        dX = torch.empty_like(X)
        dXb = ops.layernorm_bwd(ops.cast_bf16(dseq.float().contiguous()), X, mean, rstd, pk["g"], dX, False,
                                sc.g("layernorm.weight"), sc.g("layernorm.bias"))
        for i in reversed(range(len(blocks))):
            pre = f"layer.{i}."
            dXb = block_backward(dX, dXb, blocks[i], pk["layers"][i], sc, pre, vjepa_block_names(pre), rope, sq=sq)
            sq.block_done()
        pe = "embeddings.patch_embeddings." + ("proj_3d" if hasattr(self.encoder.embeddings.patch_embeddings, "proj_3d") else "proj")
        ops.colsum(dX, sc.g(pe + ".bias"))
        B, N = dX.shape[:2]
        idx = torch.arange(N, dtype=torch.int32, device=vol.device).unsqueeze(0).repeat(B, 1).contiguous()
        patches = ops.gather_patches(vol, idx, N)  # bf16 im2col rows [B*N, 4096]
        ops.linear_wgrad(dXb, patches, sc.g(pe + ".weight").view(d, -1))
        sq.finish()
        return sc.per_parameter(d)

    @torch.no_grad()
    def __call__(self, pixel_values_videos: torch.Tensor) -> torch.Tensor:
        return self.encode(self.volume(pixel_values_videos))

    def differentiable(self, pixel_values_videos: torch.Tensor) -> torch.Tensor:
        """last_hidden_state with an autograd node whose backward is `backward` above."""
        named = [(n, p) for n, p in self.encoder.named_parameters() if p.requires_grad]
        return _EncoderFunction.apply(self, pixel_values_videos, tuple(n for n, _ in named), *[p for _, p in named])


class VJepaPredictorRunner:
    """The predictor of the reference (``VJEPA2Predictor.forward``, modeling_vjepa.py:699-746, + ``VJEPA2PredictorEmbeddings``,
    :589-627, ``sort_tokens`` / ``unsort_tokens``, :658-697) on the kernels, forward AND backward, for any module laid out like
    it (our containers or the upstream class):

      context rows of the encoder output (``smbv_gather_rows_f32``) -> Linear to the predictor width (tcgen05 GEMM) | one learned
      mask token per target position -> SORTED by token position (``smbv_position_sort``: argsort, reverse argsort and the sorted
      ids in one index kernel; the sorted sequence is one row gather) -> L blocks with the sorted positions as rotary ids
      (head_dim 32 zero-padded onto the head_dim-64 attention kernels, or head_dim 64 natively) -> the TARGET rows picked straight
      out of the sorted sequence (LayerNorm is row-wise, so it runs on those rows only) -> LayerNorm -> projection back.
    """

    def __init__(self, predictor: nn.Module, config, mask_index: int = 1):  # the reference's default mask_index is 1 (:594)
        self.predictor, self.config, self.mask_index = predictor, config, mask_index
        self._packed, self._sig, self.version = None, None, 0

    def invalidate(self) -> None:
        self.version += 1

    def __deepcopy__(self, memo):
        import copy

        return VJepaPredictorRunner(copy.deepcopy(self.predictor, memo), copy.deepcopy(self.config, memo), self.mask_index)

    def check_config(self):
        c = self.config
        hd = c.pred_hidden_size // c.pred_num_attention_heads
        if hd not in (32, 64) or (hd == 32 and c.pred_num_attention_heads % 2):
            raise SmbvError("the native V-JEPA predictor implements head_dim 64 and 32 (an even number of heads): 384/12 is the reference's")
        if getattr(c, "hidden_act", "gelu") != "gelu":
            raise SmbvError("only hidden_act='gelu' (exact erf) is implemented")

    def packed(self):
        sig = (_params_signature(self.predictor), self.version)
        if self._packed is None or sig != self._sig:
            c, pr = self.config, self.predictor
            layers = [_pack_layer(l, c.pred_num_attention_heads, c.layer_norm_eps) for l in pr.layer]
            for p in layers:
                p.pad32 = p.hd == 32
            self._packed = dict(
                wemb=ops.cast_bf16(_f32(pr.embeddings.predictor_embeddings.weight)), bemb=_f32(pr.embeddings.predictor_embeddings.bias),
                token=_f32(pr.embeddings.mask_tokens)[self.mask_index % c.pred_num_mask_tokens].reshape(-1).contiguous(),
                layers=layers, g=_f32(pr.layernorm.weight), b=_f32(pr.layernorm.bias),
                wproj=ops.cast_bf16(_f32(pr.proj.weight)), bproj=_f32(pr.proj.bias))
            self._sig = sig
        return self._packed

    def _indices(self, context_mask, target_mask, dev):
        if len(context_mask) != len(target_mask):
            raise ValueError("context_mask and target_mask must be lists of the same length (reference :621 concatenates them per sample)")
        ctx = torch.cat([m.to(dev) for m in context_mask], 0).to(torch.int32).contiguous()  # [B', n_ctx]
        tgt = torch.cat([m.to(dev) for m in target_mask], 0).to(torch.int32).contiguous()   # [B', n_tgt]
        order, inv, ids, ids2 = ops.position_sort(torch.cat([ctx, tgt], 1).contiguous(), doubled=True)
        n_ctx = ctx.shape[1]
        src = torch.clamp(order, max=n_ctx)  # sorted row r comes from context row order[r], or from the mask token (row n_ctx)
        return dict(ctx=ctx, tgt=tgt, inv_ctx=inv[:, :n_ctx].contiguous(), inv_tgt=inv[:, n_ctx:].contiguous(), ids=ids, ids2=ids2,
                    src=src.contiguous(), n_ctx=n_ctx, n_tgt=tgt.shape[1], reps=len(context_mask))

    def run(self, seq: torch.Tensor, context_mask, target_mask, train: bool):
        """seq fp32 [B,N,D] (encoder output) -> predictions fp32 [B', n_tgt, D] (B' = B x number of mask pairs); `train` keeps
        what `backward` needs."""
        from .training import block_forward_train

        self.check_config()
        if train:
            self.invalidate()  # never trust version counters on a differentiable forward (see VJepaEncoderRunner.encode_train)
        pk = self.packed()
        c = self.config
        dev = seq.device
        ix = self._indices(context_mask, target_mask, dev)
        B, N, D = seq.shape
        pd, n_ctx, n_tgt = c.pred_hidden_size, ix["n_ctx"], ix["n_tgt"]
        seq = seq.float().contiguous()
        seqr = seq if ix["reps"] == 1 else seq.repeat(ix["reps"], 1, 1)
        Bp = seqr.shape[0]
        ctxb = ops.cast_bf16(ops.gather_rows(seqr, ix["ctx"]))  # [B', n_ctx, D] bf16: apply_masks(encoder_hidden_states, context_mask), :703
        S = torch.empty((Bp, n_ctx + 1, pd), dtype=torch.float32, device=dev)  # context embeddings + ONE mask-token row
        for b in range(Bp):
            ops.gemm(ctxb[b], pk["wemb"], pk["bemb"], ops.EPI_F32, out=S[b, :n_ctx])
        S[:, n_ctx] = pk["token"]
        X = ops.gather_rows(S, ix["src"])  # the concatenated [context | targets] sequence, already in sorted order (:708-709)
        gs = c.crop_size // c.patch_size
        rope = (gs, ix["ids"], min(max(gs, c.frames_per_clip // c.tubelet_size, int(N // (gs * gs)) + 1), 256), ix["ids2"])
        blocks = []
        for p in pk["layers"]:
            if train:
                X, sv = block_forward_train(X, p, rope)
                blocks.append(sv)
            else:
                _block_forward(X, p, rope)
        Xt = ops.gather_rows(X, ix["inv_tgt"])  # unsort + [:, N_ctxt:] (:732-733) as one gather; LayerNorm is row-wise
        yt, mean, rstd = ops.layernorm_fwd(Xt, pk["g"], pk["b"], c.layer_norm_eps, save_stats=True)
        pred = ops.gemm(yt, pk["wproj"], pk["bproj"], ops.EPI_F32)  # [B', n_tgt, D] fp32
        saved = (pk, ix, ctxb, blocks, Xt, yt, mean, rstd, rope, (B, N, D)) if train else None
        return pred, saved

    def backward(self, saved, dpred: torch.Tensor):
        """-> (d seq fp32 [B,N,D], {parameter name relative to the predictor: gradient})."""
        from .training import SideQueue, block_backward, vjepa_block_names

        pk, ix, ctxb, blocks, Xt, yt, mean, rstd, rope, (B, N, D) = saved
        c = self.config
        named = dict(self.predictor.named_parameters())
        sc = _ScratchGrads(named, dpred.device)
        sq = SideQueue(dpred.device)
        pd, n_ctx, n_tgt = c.pred_hidden_size, ix["n_ctx"], ix["n_tgt"]
        Bp = Xt.shape[0]
        n_tot = n_ctx + n_tgt
        dpb = ops.cast_bf16(dpred.float().contiguous())
        ops.linear_wgrad(dpb, yt, sc.g("proj.weight"))
        ops.colsum(dpb, sc.g("proj.bias"))
        dyt = ops.linear_dgrad(dpb, pk["wproj"])  # [B', n_tgt, pd] bf16
        dXt = torch.empty_like(Xt)
        ops.layernorm_bwd(dyt, Xt, mean, rstd, pk["g"], dXt, False, sc.g("layernorm.weight"), sc.g("layernorm.bias"), want_bf16=False)
        dX = ops.scatter_rows(dXt, ix["inv_tgt"], n_tot)  # only the target rows of the sorted sequence reach the output
        dXb = ops.cast_bf16(dX)
        for i in reversed(range(len(blocks))):
            pre = f"layer.{i}."
            dXb = block_backward(dX, dXb, blocks[i], pk["layers"][i], sc, pre, vjepa_block_names(pre), rope, sq=sq)
            sq.block_done()
        # sorted sequence -> its sources: the mask token (sum over every target row) and the context embeddings
        tok = torch.zeros(pd, dtype=torch.float32, device=dX.device)
        ops.colsum(ops.gather_rows(dX, ix["inv_tgt"]), tok)
        grads = None
        dctx_emb = ops.cast_bf16(ops.gather_rows(dX, ix["inv_ctx"]))  # [B', n_ctx, pd]
        ops.linear_wgrad(dctx_emb, ctxb, sc.g("embeddings.predictor_embeddings.weight"))
        ops.colsum(dctx_emb, sc.g("embeddings.predictor_embeddings.bias"))
        dctx = ops.linear_dgrad(dctx_emb, pk["wemb"], out_dtype=torch.float32)  # [B', n_ctx, D]
        dseq = ops.scatter_rows(dctx, ix["ctx"], N)  # [B', N, D]; context indices are distinct per sample
        if ix["reps"] > 1:
            dseq = dseq.view(ix["reps"], B, N, D).sum(0)
        sq.finish()
        grads = sc.per_parameter(pd)
        gm = torch.zeros_like(named["embeddings.mask_tokens"], dtype=torch.float32)
        gm[self.mask_index % c.pred_num_mask_tokens].view(-1).copy_(tok)
        grads["embeddings.mask_tokens"] = gm
        return dseq, grads


class _PredictorFunction(torch.autograd.Function):
    """predictions = predictor(encoder output, masks) as ONE autograd node (`VJepaPredictorRunner.run` / `.backward`)."""

    @staticmethod
    def forward(ctx, runner, seq, context_mask, target_mask, names, *params):
        pred, saved = runner.run(seq, context_mask, target_mask, train=True)
        ctx.runner, ctx.saved, ctx.names = runner, saved, names
        return pred

    @staticmethod
    def backward(ctx, dpred):
        dseq, grads = ctx.runner.backward(ctx.saved, dpred)
        ctx.saved = None
        return (None, dseq, None, None, None) + tuple(grads.get(n) for n in ctx.names)


def apply_masks(t: torch.Tensor, masks: List[torch.Tensor]) -> torch.Tensor:
    """reference modeling_vjepa.py:543-557: rows listed in each mask [B,K], concatenated along the batch.  Without autograd
    (inference outputs, the momentum-target rows) the gather is `smbv_gather_rows_f32`; a tensor that carries gradient
    goes through torch.gather so that autograd scatters it back."""
    out = []
    for m in masks:
        m = m.to(t.device)
        if t.is_cuda and t.dtype == torch.float32 and t.dim() == 3 and t.shape[-1] % 4 == 0 and not (torch.is_grad_enabled() and t.requires_grad):
            out.append(ops.gather_rows(t.contiguous(), m.to(torch.int32).contiguous()))
        else:
            out.append(torch.gather(t, 1, m.unsqueeze(-1).expand(-1, -1, t.size(-1))))
    return out[0] if len(out) == 1 else torch.cat(out, dim=0)


class _L1Loss(torch.autograd.Function):
    """nn.L1Loss() of the reference trainer (src/run_vjepa.py:108, :137): forward and d/d(pred) in one pass over the data."""

    @staticmethod
    def forward(ctx, pred, target):
        loss, dpred = ops.l1_loss(pred, target, want_grad=True)
        ctx.save_for_backward(dpred)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, g):
        (dpred,) = ctx.saved_tensors
        return dpred * g, None


def l1_loss(pred: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """mean |pred - target| (scalar, differentiable w.r.t. pred) on the loss kernel; fp32 CUDA tensors of equal shape."""
    pred, target = pred.float().contiguous(), target.detach().float().contiguous()
    if pred.data_ptr() % 16:
        pred = pred.clone()
    if target.data_ptr() % 16:
        target = target.clone()
    return _L1Loss.apply(pred, target) if (torch.is_grad_enabled() and pred.requires_grad) else ops.l1_loss(pred, target).reshape(())


class B200VJEPA2Model(_PretrainedIO, nn.Module):
    """Reference ``VJEPA2Model`` (modeling_vjepa.py:1057-1153) with the encoder on the CUDA kernels."""

    base_model_prefix = "vjepa2"
    main_input_name = "pixel_values_videos"

    def __init__(self, config, with_predictor: bool = True):
        super().__init__()
        self.config = config
        d = config.hidden_size
        if d % config.num_attention_heads != 0:
            raise ValueError(f"The hidden size {(d,)} is not a multiple of the number of attention heads {config.num_attention_heads}.")
        self.encoder = _VJepaEncoder(config)
        _init_weights(self.encoder, getattr(config, "initializer_range", 0.02))
        self.predictor = None
        self._pred_runner = None
        if with_predictor:  # same parameter names / shapes as the reference's (and upstream's) VJEPA2Predictor
            self.predictor = _VJepaPredictor(config)
            _init_weights(self.predictor, getattr(config, "initializer_range", 0.02))
            if not getattr(config, "pred_zero_init_mask_tokens", True):  # reference :1031-1035
                nn.init.trunc_normal_(self.predictor.embeddings.mask_tokens, std=getattr(config, "initializer_range", 0.02))
            self._pred_runner = VJepaPredictorRunner(self.predictor, config)
        self._runner = VJepaEncoderRunner(self.encoder, config)
        self._arena = None

    @classmethod
    def _config_from_dir(cls, path):
        from transformers import VJEPA2Config

        return VJEPA2Config.from_pretrained(path)

    def _rename_checkpoint_keys(self, sd: dict) -> dict:
        # upstream transformers names the tubelet convolution `proj`; the reference (in_chans-configurable copy) `proj_3d`
        pre = "encoder.embeddings.patch_embeddings."
        return {(pre + "proj_3d" + k[len(pre) + 4:] if k.startswith(pre + "proj.") else k): v for k, v in sd.items()}

    def load_state_dict(self, state_dict, strict: bool = True, assign: bool = False):
        return super().load_state_dict(self._rename_checkpoint_keys(state_dict), strict=strict, assign=assign)

    def get_input_embeddings(self):
        return self.encoder.embeddings.patch_embeddings

    def packed(self):
        return self._runner.packed()

    def _volume(self, pixel_values_videos: torch.Tensor) -> torch.Tensor:
        return self._runner.volume(pixel_values_videos)

    def encode(self, vol: torch.Tensor) -> torch.Tensor:
        return self._runner.encode(vol)

    def forward(self, pixel_values_videos: torch.Tensor, context_head_mask=None, context_mask: Optional[List[torch.Tensor]] = None,
                target_head_mask=None, target_mask: Optional[List[torch.Tensor]] = None, skip_predictor: bool = False,
                output_attentions: Optional[bool] = None, output_hidden_states: Optional[bool] = None, **kwargs):
        if context_head_mask is not None or target_head_mask is not None:
            raise ValueError("head masks are not supported by the fused attention kernel")
        if output_attentions:
            raise ValueError("output_attentions is not supported by the fused attention kernel")
        if not skip_predictor and self.predictor is None:
            raise SmbvError("this model was built without the predictor (with_predictor=False): pass skip_predictor=True")
        grad = torch.is_grad_enabled() and any(p.requires_grad for p in self.encoder.parameters())
        with torch.enable_grad() if grad else torch.no_grad():
            # training: the encoder and the predictor are ONE autograd node each, with hand-written backward passes
            seq = self._runner.differentiable(pixel_values_videos) if grad else self._runner(pixel_values_videos)
            B, N = seq.shape[:2]
            if context_mask is None and target_mask is None:  # reference :1120-1124
                ar = torch.arange(N, device=seq.device).unsqueeze(0).repeat((B, 1))
                context_mask, target_mask = [ar], [ar]
            pred = None
            if not skip_predictor:
                cm, tm = [m.to(seq.device) for m in context_mask], [m.to(seq.device) for m in target_mask]
                if torch.is_grad_enabled() and (seq.requires_grad or any(p.requires_grad for p in self.predictor.parameters())):
                    named = [(n, p) for n, p in self.predictor.named_parameters() if p.requires_grad]
                    ph = _PredictorFunction.apply(self._pred_runner, seq, cm, tm, tuple(n for n, _ in named), *[p for _, p in named])
                else:
                    ph = self._pred_runner.run(seq, cm, tm, train=False)[0]
                pred = VJEPA2WithMaskedInputPredictorOutput(last_hidden_state=ph, target_hidden_state=apply_masks(seq, target_mask))
            return VJEPA2WithMaskedInputModelOutput(last_hidden_state=seq, masked_hidden_state=apply_masks(seq, context_mask),
                                                    target_hidden_state=apply_masks(seq, target_mask), predictor_output=pred)

    def invalidate_packed(self) -> None:
        """Call after an optimiser that updates the parameters without bumping torch's version counters (FusedAdamW does
        it itself)."""
        self._runner.invalidate()
        if self._pred_runner is not None:
            self._pred_runner.invalidate()

    def get_vision_features(self, pixel_values_videos) -> torch.Tensor:
        """reference :1151-1153 (`forward(x).last_hidden_state`; the reference also runs its predictor there and throws the
        result away — here only the encoder runs)."""
        return self._runner(pixel_values_videos)
