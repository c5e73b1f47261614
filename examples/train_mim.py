"""MIM pre-training loop on the B200 path — the role of the reference's `src/run_mim.py` + HF `Trainer` for this path.

    python examples/train_mim.py --synthetic 8 --steps 20                      # one GPU, synthetic CT-like volumes
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 examples/train_mim.py --volumes /data/*.npy

One process per GPU.  Per step: raw volume (fp32 or int16 HU, [X,Y,Z]) -> `VolumePreprocessor` (scale/pad/crop/permute on the
GPU) -> `MaskGenerator.device_batch` (reference RNG stream, index lists on the GPU) -> `DataParallelStep` (CUDA forward +
backward, bucketed bf16 all-reduce overlapped with backward) -> `FusedAdamW` (clip 1.0 + AdamW + bf16 operand refresh) with
the cosine/warm-up schedule of scripts/training/run_mim.sh:17-21.  Checkpoints are `save_pretrained` directories that the
reference's `VideoMAEForPreTraining.from_pretrained` loads.
"""
from __future__ import annotations

import argparse
import functools
import glob
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--volumes", nargs="*", default=[], help=".npy files holding one resampled volume [X,Y,Z] each (fp32 or int16 HU)")
    ap.add_argument("--synthetic", type=int, default=0, help="use this many synthetic int16 volumes instead of files")
    ap.add_argument("--model_name_or_path", default=None)
    ap.add_argument("--image_size", type=int, default=512)
    ap.add_argument("--depth", type=int, default=320)
    ap.add_argument("--mask_patch_size", type=int, default=32)
    ap.add_argument("--mask_ratio", type=float, default=0.65)
    ap.add_argument("--batch", type=int, default=1, help="volumes per GPU per step")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--learning_rate", type=float, default=5e-5)
    ap.add_argument("--weight_decay", type=float, default=0.01)
    ap.add_argument("--max_grad_norm", type=float, default=1.0)
    ap.add_argument("--warmup_ratio", type=float, default=0.01)
    ap.add_argument("--seed", type=int, default=42)
    ap.add_argument("--output_dir", default=None)
    ap.add_argument("--cuda_graph", action="store_true", help="replay forward + backward (+ all-reduce) from a CUDA graph (DataParallelStep(cuda_graph=True))")
    ap.add_argument("--config_overrides", default="", help="k=v,k=v VideoMAEConfig overrides (e.g. a small test model)")
    args = ap.parse_args(argv)

    from transformers import VideoMAEConfig

    from smb_vision_b200.data import MaskGenerator, VolumePreprocessor
    from smb_vision_b200.distributed import shard_volumes
    from smb_vision_b200.modeling import B200VideoMAEForPreTraining
    from smb_vision_b200.optim import FusedAdamW, cosine_with_warmup
    from smb_vision_b200.training import DataParallelStep

    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1 and not dist.is_initialized():
        dist.init_process_group("nccl", device_id=dev)

    if args.model_name_or_path:
        config = VideoMAEConfig.from_pretrained(args.model_name_or_path)
    else:
        config = VideoMAEConfig()
    config.update({k: type(getattr(config, k))(v) for k, v in (kv.split("=") for kv in args.config_overrides.split(",") if kv)})
    config.update(dict(image_size=args.image_size, patch_size=16, num_channels=1, num_frames=args.depth, tubelet_size=16))  # run_mim.py:322-330
    torch.manual_seed(args.seed)  # same initial weights on every rank
    model = (B200VideoMAEForPreTraining.from_pretrained(args.model_name_or_path, config=config) if args.model_name_or_path
             else B200VideoMAEForPreTraining(config)).to(dev).train()

    files = sorted(f for pat in args.volumes for f in glob.glob(pat))
    if args.synthetic:
        g = torch.Generator().manual_seed(args.seed)
        raws = [torch.randint(-1100, 1500, (args.image_size, args.image_size, args.depth), generator=g, dtype=torch.int16).pin_memory()
                for _ in range(args.synthetic)]
    else:
        raws = [torch.from_numpy(np.load(f)) for f in files]
    raws = shard_volumes(raws, rank, world) or raws[:1]
    np.random.seed(args.seed + rank)  # masks differ per rank

    sched = functools.partial(cosine_with_warmup, base_lr=args.learning_rate, warmup_steps=int(args.warmup_ratio * args.steps + 0.999),
                              total_steps=args.steps)
    opt = FusedAdamW(model, lr=args.learning_rate, weight_decay=args.weight_decay, max_grad_norm=args.max_grad_norm, lr_schedule=sched)
    dp = DataParallelStep(model, optimizer=opt, cuda_graph=args.cuda_graph)
    prep = VolumePreprocessor(args.image_size, args.depth, device=dev)
    masks = MaskGenerator(args.image_size, args.depth, args.mask_patch_size, 16, args.mask_ratio)

    losses = []
    for step in range(args.steps):
        batch = [raws[(step * args.batch + i) % len(raws)] for i in range(args.batch)]
        vol = prep.batch(batch).view(args.batch, args.depth, args.image_size, args.image_size)
        loss, _ = dp.step(vol, masks.device_batch(args.batch, dev))
        losses.append(loss.clone())  # (graph mode returns a static tensor that the next step overwrites)
        if rank == 0 and (step % 10 == 0 or step == args.steps - 1):
            print(f"step {step} loss {float(loss):.6f} lr {opt.current_lr():.3e} grad_norm {float(opt.grad_norm()):.4f}", flush=True)
    if args.output_dir and rank == 0:
        model.save_pretrained(args.output_dir)
        torch.save(opt.state_dict(), os.path.join(args.output_dir, "optimizer.pt"))
    dp.release_graphs()  # before any process-group teardown: the graphs hold captured all-reduces
    if world > 1:
        dist.barrier()
    return [float(x) for x in losses]


if __name__ == "__main__":
    main()
