"""V-JEPA2-3D encoder on the sm_100a kernels: the path the reference runs for embedding extraction
and for its momentum TARGET encoder — ``VJEPA2Model.forward(pixel_values_videos, context_mask, target_mask,
skip_predictor=True)`` (reference src/models/vjepa/modeling_vjepa.py:1071-1149, called under ``torch.no_grad()`` at
src/run_vjepa.py:128-135) and ``get_vision_features`` (:1151-1153).

Same constructor (a ``VJEPA2Config`` as ``src/run_vjepa.py:220-232`` builds it: ``in_chans=1``, ``tubelet_size = patch_size
= 16``), same parameter names and shapes (``encoder.embeddings.patch_embeddings.proj_3d``, ``encoder.layer.N.{norm1,
attention.{query,key,value,proj}, norm2, mlp.{fc1,fc2}}``, ``encoder.layernorm``, ``predictor.*``), same output class.

Encoder forward = tubelet embedding (implicit-GEMM kernel, no position table) -> L x [LayerNorm -> fused QKV GEMM with all
three biases, head-major -> ``smbv_rope3d`` in place on Q and K -> tcgen05 flash attention -> proj + residual -> LayerNorm
-> fc1 + GELU -> fc2 + residual] -> final LayerNorm.  The predictor (12 x 384/12, head_dim 32) is the upstream module
driven through the attention plug-in (`attention_interface.py`); it is only built when transformers provides it.
With gradients enabled the encoder is ONE autograd node with a hand-written backward (`VJepaEncoderRunner.backward`: the
VideoMAE block backward of `training.py` + the transposed rotary map + K-bias gradient), so the online model of
`examples/train_vjepa.py --native_online` trains with its encoder entirely on the kernels.
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
        from .training import block_backward, vjepa_block_names

        pk, blocks, X, mean, rstd, rope = saved
        named = dict(self.encoder.named_parameters())
        sc = _ScratchGrads(named, vol.device)
        d = self.config.hidden_size
        dX = torch.empty_like(X)
        dXb = ops.layernorm_bwd(ops.cast_bf16(dseq.float().contiguous()), X, mean, rstd, pk["g"], dX, False,
                                sc.g("layernorm.weight"), sc.g("layernorm.bias"))
        for i in reversed(range(len(blocks))):
            pre = f"layer.{i}."
            dXb = block_backward(dX, dXb, blocks[i], pk["layers"][i], sc, pre, vjepa_block_names(pre), rope)
        pe = "embeddings.patch_embeddings." + ("proj_3d" if hasattr(self.encoder.embeddings.patch_embeddings, "proj_3d") else "proj")
        ops.colsum(dX, sc.g(pe + ".bias"))
        B, N = dX.shape[:2]
        idx = torch.arange(N, dtype=torch.int32, device=vol.device).unsqueeze(0).repeat(B, 1).contiguous()
        patches = ops.gather_patches(vol, idx, N)  # bf16 im2col rows [B*N, 4096]
        ops.linear_wgrad(dXb, patches, sc.g(pe + ".weight").view(d, -1))
        return sc.per_parameter(d)

    @torch.no_grad()
    def __call__(self, pixel_values_videos: torch.Tensor) -> torch.Tensor:
        return self.encode(self.volume(pixel_values_videos))

    def differentiable(self, pixel_values_videos: torch.Tensor) -> torch.Tensor:
        """last_hidden_state with an autograd node whose backward is `backward` above."""
        named = [(n, p) for n, p in self.encoder.named_parameters() if p.requires_grad]
        return _EncoderFunction.apply(self, pixel_values_videos, tuple(n for n, _ in named), *[p for _, p in named])


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
        if with_predictor:
            try:
                from transformers.models.vjepa2.modeling_vjepa2 import VJEPA2Predictor
            except Exception:
                VJEPA2Predictor = None
            if VJEPA2Predictor is not None:
                from . import attention_interface

                attention_interface.register()
                config._attn_implementation = attention_interface.NAME
                self.predictor = VJEPA2Predictor(config)
                _init_weights(self.predictor, getattr(config, "initializer_range", 0.02))
                if not getattr(config, "pred_zero_init_mask_tokens", True):  # reference :1031-1035
                    nn.init.trunc_normal_(self.predictor.embeddings.mask_tokens, std=getattr(config, "initializer_range", 0.02))
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
            raise SmbvError("this model was built without the predictor (transformers' VJEPA2Predictor not available): pass skip_predictor=True")
        grad = torch.is_grad_enabled() and any(p.requires_grad for p in self.encoder.parameters())
        with torch.enable_grad() if grad else torch.no_grad():
            # training: the encoder is one autograd node with a hand-written backward; the predictor (torch, attention
            # through the plug-in) and the mask gathers are ordinary autograd on top of it
            seq = self._runner.differentiable(pixel_values_videos) if grad else self._runner(pixel_values_videos)
            B, N = seq.shape[:2]
            if context_mask is None and target_mask is None:  # reference :1120-1124
                ar = torch.arange(N, device=seq.device).unsqueeze(0).repeat((B, 1))
                context_mask, target_mask = [ar], [ar]
            pred = None
            if not skip_predictor:  # upstream predictor, attention through the plug-in (head_dim 32 kernels)
                po = self.predictor(encoder_hidden_states=seq, context_mask=[m.to(seq.device) for m in context_mask],
                                    target_mask=[m.to(seq.device) for m in target_mask])
                pred = VJEPA2WithMaskedInputPredictorOutput(last_hidden_state=po.last_hidden_state,
                                                            target_hidden_state=apply_masks(seq, target_mask))
            return VJEPA2WithMaskedInputModelOutput(last_hidden_state=seq, masked_hidden_state=apply_masks(seq, context_mask),
                                                    target_hidden_state=apply_masks(seq, target_mask), predictor_output=pred)

    def invalidate_packed(self) -> None:
        """Call after an optimiser that updates the parameters without bumping torch's version counters (FusedAdamW does
        it itself)."""
        self._runner.invalidate()

    def get_vision_features(self, pixel_values_videos) -> torch.Tensor:
        """reference :1151-1153 (`forward(x).last_hidden_state`; the reference also runs its predictor there and throws the
        result away — here only the encoder runs)."""
        return self._runner(pixel_values_videos)
