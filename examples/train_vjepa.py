"""V-JEPA 3D training step with the B200 kernels behind the reference's own plug-in points (first slice).

The role of `src/run_vjepa.py` (`VJEPATrainer.compute_loss`, :101-137, + `MomentumEncoder`, :87-99) for one process per GPU:
the UNMODIFIED `transformers.VJEPA2Model` (what the reference's vendored `modeling_vjepa.py` tracks) runs with
`config._attn_implementation = "b200_tcgen05"`, so every RoPE attention — encoder at head_dim 64 on the tcgen05 kernels, predictor
at head_dim 32 on the small-head kernels — goes through our forward AND backward kernels; gradients accumulate in a flat arena
(`FusedAdamW.grad_arena()`), `smbv_sumsq_f32` + `smbv_adamw_step` do clip + AdamW, and `smbv_ema_update` moves the target encoder.
The momentum TARGET encoder's forward — half of the encoder forward work of a step — runs on the native V-JEPA encoder
(`smb_vision_b200/vjepa.py`: rotary kernel, fused QKV with K bias, tcgen05 GEMMs + attention) straight from the momentum weights;
the online model's linear layers / LayerNorm / RoPE still run in torch (bf16 autocast) — unless `--native_online` swaps in
`B200VJEPA2Model`, whose encoder AND predictor are one autograd node each with hand-written backward passes (no torch math left
in the step: `smb_vision_b200/vjepa.py`, `VJepaEncoderRunner` / `VJepaPredictorRunner`).

    python examples/train_vjepa.py --steps 10 [--native_online]
"""
from __future__ import annotations

import argparse
import os
import sys

import torch
import torch.distributed

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def vjepa_fwd_bwd(model, target, grads, x, context_mask, target_mask, native_target: bool = True):
    """Forward (online model + momentum target) + L1 + backward + data-parallel gradient mean of one step; no optimiser state
    is touched, so this part can be captured in a CUDA graph (`GraphedVJEPAStep`)."""
    from smb_vision_b200.vjepa import apply_masks, l1_loss  # smbv_gather_rows_f32 / smbv_l1_loss_f32

    grads.zero()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        out = model(pixel_values_videos=x, context_mask=context_mask, target_mask=target_mask)
        predicted = out.predictor_output.last_hidden_state
        with torch.no_grad():
            if native_target:
                tgt = apply_masks(target.encode(x), target_mask)
            else:
                t_out = target.model(pixel_values_videos=x, context_mask=context_mask, target_mask=target_mask, skip_predictor=True)
                tgt = apply_masks(t_out.last_hidden_state, target_mask)
    loss = l1_loss(predicted, tgt)  # nn.L1Loss(), forward + gradient in one pass
    loss.backward()          # autograd accumulates straight into the flat gradient arena
    grads.all_reduce()       # data-parallel mean over the ranks (accelerate's DDP in the reference); no-op for one process
    return loss.detach()


def vjepa_step(model, target, opt, grads, x, context_mask, target_mask, native_target: bool = True):
    """One optimisation step; returns the L1 loss (reference src/run_vjepa.py:108-137).  native_target: the momentum target
    encoder's forward runs on the native encoder kernels (`EmaTarget.encode`, head_dim 64) instead of torch + plug-in."""
    loss = vjepa_fwd_bwd(model, target, grads, x, context_mask, target_mask, native_target)
    opt.step(grads)          # clip_grad_norm_ + AdamW, one pass
    target.update()          # momentum update of the target encoder, one pass
    return loss


class GraphedVJEPAStep:
    """The same step with forward + backward replayed from a CUDA graph (one graph per mask-shape signature: the block masks of
    `VJEPAMaskGenerator` keep their token counts for a fixed geometry, only the indices change).  The step is ~7 700 launches of
    mostly small kernels, i.e. launch-bound when issued one by one; clip + AdamW and the EMA update (step-dependent scalars) stay
    outside the graph.  The returned loss is a static tensor, overwritten by the next step."""

    def __init__(self, model, target, opt, grads, native_target: bool = True):
        self.model, self.target, self.opt, self.grads, self.native_target = model, target, opt, grads, native_target
        self._graphs = {}

    def _capture(self, key, x, ctx, tgt):
        sx, sc, st = x.clone(), [m.clone() for m in ctx], [m.clone() for m in tgt]
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            for _ in range(2):
                vjepa_fwd_bwd(self.model, self.target, self.grads, sx, sc, st, self.native_target)
        cur.wait_stream(side)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            loss = vjepa_fwd_bwd(self.model, self.target, self.grads, sx, sc, st, self.native_target)
        self._graphs[key] = (g, sx, sc, st, loss)

    def __call__(self, x, context_mask, target_mask):
        key = (tuple(x.shape),) + tuple(tuple(m.shape) for m in context_mask) + tuple(tuple(m.shape) for m in target_mask)
        if key not in self._graphs:
            self._capture(key, x, context_mask, target_mask)
        g, sx, sc, st, loss = self._graphs[key]
        for src, dst in [(x, sx)] + list(zip(context_mask, sc)) + list(zip(target_mask, st)):
            if src.data_ptr() != dst.data_ptr():
                dst.copy_(src, non_blocking=True)
        g.replay()
        self.opt.step(self.grads)
        self.target.update()
        return loss


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--image_size", type=int, default=64)
    ap.add_argument("--depth", type=int, default=64)
    ap.add_argument("--hidden_size", type=int, default=128)
    ap.add_argument("--heads", type=int, default=2)
    ap.add_argument("--layers", type=int, default=2)
    ap.add_argument("--pred_hidden_size", type=int, default=64)
    ap.add_argument("--pred_heads", type=int, default=2)
    ap.add_argument("--pred_layers", type=int, default=2)
    ap.add_argument("--batch", type=int, default=2)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--learning_rate", type=float, default=1e-3)
    ap.add_argument("--attn", default="b200_tcgen05")
    ap.add_argument("--native_online", action="store_true",
                    help="online model = B200VJEPA2Model: encoder and predictor forward AND backward on the kernels")
    ap.add_argument("--torch_target", action="store_true", help="run the target encoder in torch through the plug-in instead of natively")
    args = ap.parse_args(argv)

    import transformers

    import smb_vision_b200.attention_interface as ai
    from smb_vision_b200.optim import EmaTarget, FusedAdamW

    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0)))
    torch.cuda.set_device(dev)
    if int(os.environ.get("WORLD_SIZE", "1")) > 1 and not torch.distributed.is_initialized():  # torchrun: one process per GPU
        torch.distributed.init_process_group("nccl", device_id=dev)
    rank = torch.distributed.get_rank() if torch.distributed.is_initialized() else 0
    ai.register()
    c = transformers.VJEPA2Config(patch_size=16, crop_size=args.image_size, frames_per_clip=args.depth, tubelet_size=16, in_chans=1,
                                  hidden_size=args.hidden_size, num_attention_heads=args.heads, num_hidden_layers=args.layers,
                                  pred_hidden_size=args.pred_hidden_size, pred_num_attention_heads=args.pred_heads,
                                  pred_num_hidden_layers=args.pred_layers, pred_num_mask_tokens=2)  # src/run_vjepa.py:222-234
    c._attn_implementation = args.attn
    torch.manual_seed(0)
    if args.native_online:
        from smb_vision_b200.vjepa import B200VJEPA2Model

        model = B200VJEPA2Model(c).to(dev).train()
    else:
        model = transformers.VJEPA2Model(c).to(dev).train()
    opt = FusedAdamW(model, lr=args.learning_rate, weight_decay=0.01, max_grad_norm=1.0)
    grads = opt.grad_arena()
    target = EmaTarget(model, momentum=0.99925)
    from smb_vision_b200.data import VJEPAMaskGenerator, vjepa_collate_fn

    g = torch.Generator().manual_seed(1 + rank)  # every rank draws its own volumes (and masks: torch's global stream below)
    torch.manual_seed(100 + rank)
    # block masks as the "vjepa" transform preset draws them (src/dataloader/transforms.py:257-263), in token order (frames, rows, columns)
    masks = VJEPAMaskGenerator(input_size=(args.depth, args.image_size, args.image_size), patch_size=(16, 16, 16), num_blocks=3)
    losses = []
    for step in range(args.steps):
        examples = []
        while len(examples) < args.batch:
            ex = masks({"image": torch.rand(args.depth, 1, args.image_size, args.image_size, generator=g)})
            if ex["context_mask"].numel() >= 2 and ex["target_mask"].numel() >= 2:  # tiny grids: three blocks can cover everything
                examples.append(ex)
        batch = vjepa_collate_fn(examples)  # one example's masks shared across the batch (src/run_vjepa.py:144-160)
        x = batch["pixel_values_videos"].to(dev)
        ctx, tgt = [m.to(dev) for m in batch["context_mask"]], [m.to(dev) for m in batch["target_mask"]]
        losses.append(float(vjepa_step(model, target, opt, grads, x, ctx, tgt, native_target=not args.torch_target)))
        if step % 5 == 0 or step == args.steps - 1:
            print(f"step {step} loss {losses[-1]:.5f} grad_norm {float(opt.grad_norm()):.4f}", flush=True)
    return losses


if __name__ == "__main__":
    main()
