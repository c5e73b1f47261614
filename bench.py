"""Headline benchmark: volumes/sec at 512x512x320 (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

At N=1 the workload is BASELINE.json configs[1]: smb-vision-base embedding inference, bf16 tensor-core compute,
random-init weights, synthetic 512x512x320 volumes — one "step" = `model.videomae(x)` on one volume per GPU
(reference src/run_inference.py:78-86).  Under torchrun (N>1) every rank processes its own volumes
(volume sharding, no collective — the reference's run_inspect.py:206-241 strategy); `value` = volumes all ranks
processed / max-over-ranks device time.

`--impl reference` times the reference's own CPU path — upstream `transformers.VideoMAEModel`, the class the reference
imports, fp32, sdpa backend, all host threads — on a bounded sample of the same workload (see run_reference).
"""
from __future__ import annotations

import argparse
import contextlib
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "volumes/sec at 512x512x320: embedding inference (MIM train step reported beside it)"
UNIT = "volumes/s"
WORKLOAD = "smb-vision-base embedding inference, 512x512x320 (20480 tokens), batch 1 volume/GPU, bf16 operands fp32 accumulate"
BASE = {}  # OracleConfig defaults == smb-vision-base at 512x512x320

# algorithmic FLOPs per volume
N_TOK, D, HEADS, LAYERS, MLP = 20480, 768, 12, 12, 3072
ATTN_FLOPS_PER_LAUNCH = 4.0 * N_TOK * N_TOK * 64 * HEADS  # 1.2885 TFLOP
ATTN_DRAM_BYTES_PER_LAUNCH = 94.49e6 + 17.46e6  # ncu --set full, profiles/r01_ncu_full.md (refreshed when the kernel changes)
EMBED_FLOPS = 2.0 * N_TOK * 4096 * D + LAYERS * (2.0 * N_TOK * D * (3 * D + D + 2 * MLP) + ATTN_FLOPS_PER_LAUNCH)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        j = json.load(open(p))
        return dict(hbm=j["hbm_gbs"], tf_burst=j["bf16_tflops"], tf_sust=j.get("bf16_tflops_sustained", j["bf16_tflops"]), src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx, self.proc, self.lines = gpu_index, None, []

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True)
            self.th.start()
        except Exception:
            self.proc = None
        return self

    def __exit__(self, *a):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=3)
            except Exception:
                self.proc.kill()
            self.th.join(timeout=2)

    def summary(self):
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx = max(mx, float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=[], samples=0)
        busy = [v for v in sm if v > 0.5 * max(sm)] or sm
        return dict(sm_mhz=statistics.median(busy), sm_max_mhz=mx, reasons=sorted(reasons), samples=len(sm))


# ------------------------------------------------------------------------------------------
# CPU arm: oracle port of the reference path on the host cores (bounded sample)
# ------------------------------------------------------------------------------------------
def cpu_embed_sample(layers_sampled: int = 1, repeats: int = 1):
    """Times patch-embed + `layers_sampled` of the 12 encoder layers at the full 20480 tokens (fp32, sdpa backend,
    all host threads) and extrapolates linearly in the layer count.  Returns (seconds_per_volume, cores, sample)."""
    import torch
    from oracle import videomae_oracle as vo

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    vo.ATTN_IMPL = "sdpa"
    cfg = vo.OracleConfig(**BASE)
    g = torch.Generator().manual_seed(1234)
    shapes = vo.param_shapes(cfg)
    sd = {}
    for k, shp in shapes.items():
        if k.startswith("videomae.embeddings") or any(k.startswith(f"videomae.encoder.layer.{i}.") for i in range(layers_sampled)):
            sd[k] = (torch.ones(shp) if ("layernorm" in k and k.endswith("weight")) else 0.02 * torch.randn(shp, generator=g)).float()
    x = vo.synthetic_volume(cfg, 1, 7)
    best_e, best_l = 1e30, 1e30
    with torch.no_grad():
        for _ in range(repeats):
            t0 = time.perf_counter()
            h = vo.embed(sd, cfg, x, None)
            t1 = time.perf_counter()
            for i in range(layers_sampled):
                h = vo._layer(h, sd, f"videomae.encoder.layer.{i}.", cfg.num_attention_heads, cfg.layer_norm_eps)
            t2 = time.perf_counter()
            best_e, best_l = min(best_e, t1 - t0), min(best_l, (t2 - t1) / layers_sampled)
    vo.ATTN_IMPL = "eager"
    total = best_e + cfg.num_hidden_layers * best_l
    sample = (f"oracle port of modeling_videomae.py (fp32, sdpa backend): patch-embed + {layers_sampled} of 12 encoder layers at the full "
              f"20480 tokens ({best_e:.2f}s + {best_l:.2f}s/layer), extrapolated x12 layers")
    return total, cores, sample


def run_reference(args, rank):
    """`--impl reference`: the reference's OWN model class — `transformers.VideoMAEModel`, which is what src/run_inference.py:12
    / src/run_mim.py:19-20 import — through its public API `model(x).last_hidden_state` (run_inference.py:78-86), unmodified,
    fp32, sdpa backend, all host threads, random init, on the same synthetic 512x512x320 volume.  One step = one forward of
    the real class with L of the 12 encoder layers (L sized from a one-layer calibration so that warm-up + K steps finish in
    a few minutes), extrapolated linearly in the layer count; L = 12 (no extrapolation) when the budget allows."""
    if rank != 0:
        return
    import torch
    import transformers

    from __graft_entry__ import hf_config
    from oracle import videomae_oracle as vo

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    ocfg = vo.OracleConfig(**BASE)
    cfgd = {k: getattr(ocfg, k) for k in ocfg.__dataclass_fields__}
    x = vo.synthetic_volume(ocfg, 1, 7)

    hc = hf_config(cfgd)
    hc._attn_implementation = "sdpa"
    torch.manual_seed(1234)
    model = transformers.VideoMAEModel(hc).eval()  # (its __init__ builds the sin-cos table in pure Python: ~40 s at this size)
    layers = list(model.encoder.layer)
    warm = 1 if args.warmup >= 1 else 0
    with torch.no_grad():
        t0 = time.perf_counter()
        h = model.embeddings(x, None)  # calibration: embeddings alone, then one encoder layer
        t_embed = time.perf_counter() - t0
        t0 = time.perf_counter()
        layers[0](h)
        t_layer = max(time.perf_counter() - t0, 1e-3)
        del h
        budget = 90.0
        L = int(max(1, min(ocfg.num_hidden_layers, (budget / (args.steps + warm) - t_embed) // t_layer)))
        if L < len(layers):  # bounded sample: the same module stack, cut after L layers (what num_hidden_layers = L builds)
            model.encoder.layer = torch.nn.ModuleList(layers[:L])
        for _ in range(warm):
            model(x)
        times = []
        for _ in range(args.steps):
            t0 = time.perf_counter()
            y = model(x).last_hidden_state
            times.append(time.perf_counter() - t0)
        assert tuple(y.shape) == (1, N_TOK, D)
    t_L = sum(times) / len(times)
    t = t_embed + ocfg.num_hidden_layers * max(t_L - t_embed, 1e-6) / L
    del model, layers
    mim = None
    if not args.no_mim:
        mim = reference_mim_sample(hf_config(cfgd), x, cores)
    sample = (f"transformers.VideoMAEModel (the class the reference imports), fp32, sdpa, {cores} threads, full 512x512x320 volume: forward with "
              f"{L} of 12 encoder layers = {t_L:.2f}s/step (embeddings {t_embed:.2f}s)"
              + ("" if L == ocfg.num_hidden_layers else ", extrapolated linearly to 12 layers"))
    line = {
        "impl": "reference", "metric": METRIC, "value": 1.0 / t, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": warm,
        "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "arm": "the reference's own CPU path (upstream VideoMAEModel, fp32, sdpa) on the host cores, rank 0 only, bounded sample"},
        "cpu_baseline": {"value": 1.0 / t, "unit": UNIT, "cores": cores, "kind": "reference", "sample": sample},
        "e2e": {"value": 1.0 / t, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    if mim:
        line["mim"] = mim
    print(json.dumps(line), flush=True)


def reference_mim_sample(hc, x, cores):
    """The reference's MIM training step on the host cores: upstream `transformers.VideoMAEForPreTraining` (what src/run_mim.py:19-20
    imports) `model(x, mask).loss.backward()` at the full 512x512x320 size, fp32, sdpa.  Bounded sample: the real module stacks cut
    to (1,1), (2,1) and (1,2) (encoder, decoder) layers; the per-layer costs are solved from the three timings and extrapolated to
    12 + 4 layers."""
    import numpy as np
    import torch
    import transformers

    from oracle.mim_mask import OracleMaskGenerator

    hc._attn_implementation = "sdpa"
    torch.manual_seed(1234)
    model = transformers.VideoMAEForPreTraining(hc).train()
    enc, dec = list(model.videomae.encoder.layer), list(model.decoder.decoder_layers)
    np.random.seed(0)
    mask = torch.from_numpy(OracleMaskGenerator(512, 320, 32, 16, 0.65)()).unsqueeze(0)

    def run(le, ld):
        model.videomae.encoder.layer = torch.nn.ModuleList(enc[:le])
        model.decoder.decoder_layers = torch.nn.ModuleList(dec[:ld])
        model.zero_grad(set_to_none=True)
        t0 = time.perf_counter()
        model(x, mask).loss.backward()
        return time.perf_counter() - t0

    t11, t21, t12 = run(1, 1), run(2, 1), run(1, 2)
    te, td = max(t21 - t11, 1e-3), max(t12 - t11, 1e-3)
    base = max(t11 - te - td, 0.0)
    total = base + len(enc) * te + len(dec) * td
    return {"train_step_s": total, "volumes_per_s": 1.0 / total, "cores": cores, "kind": "reference",
            "sample": (f"transformers.VideoMAEForPreTraining forward + backward (fp32, sdpa, {cores} threads, full volume, 65% masked): "
                       f"(enc,dec) layers (1,1) {t11:.1f}s, (2,1) {t21:.1f}s, (1,2) {t12:.1f}s -> {te:.1f}s per encoder layer, {td:.1f}s per decoder layer, "
                       f"{base:.1f}s embeddings/head/loss; extrapolated to {len(enc)} + {len(dec)} layers; no optimiser step")}


# ------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-mim", action="store_true")
    ap.add_argument("--no-vjepa", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import numpy as np
    import torch
    import torch.distributed as dist

    from __graft_entry__ import hf_config
    from oracle import videomae_oracle as vo  # synthetic inputs only; never on the timed path
    from oracle.mim_mask import OracleMaskGenerator
    from smb_vision_b200 import _lib
    from smb_vision_b200.modeling import B200VideoMAEForPreTraining

    assert torch.cuda.is_available(), "bench.py needs a GPU (use --impl reference for the CPU arm)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    args.warmup = max(args.warmup, 3)

    ocfg = vo.OracleConfig(**BASE)
    cfgd = {k: getattr(ocfg, k) for k in ocfg.__dataclass_fields__}
    torch.manual_seed(1234)
    model = B200VideoMAEForPreTraining(hf_config(cfgd)).to(dev).eval()
    x_host = vo.synthetic_volume(ocfg, 1, 7 + rank).pin_memory()  # [1,320,1,512,512] fp32, 335.5 MB (> 126 MB L2)
    x_dev = x_host.to(dev)
    emb_host = torch.empty((1, N_TOK, D), dtype=torch.float32).pin_memory()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- per-kernel CUDA-event timing of the dominant kernel (flash attention) inside the timed region ----
    attn_events = []

    @contextlib.contextmanager
    def hook(name):
        if name == "smbv_flash_attn_fwd_ex":
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            yield
            e1.record()
            attn_events.append((e0, e1))
        else:
            yield

    def timed(fn, steps, with_hook=False):
        # the sampler starts before the warm-up (same load) so that nvidia-smi is already producing samples when the
        # timed region begins; only samples under load are summarised
        import gc

        with ClockSampler(local_rank) as cs:
            for _ in range(args.warmup):
                fn()
            was_enabled = gc.isenabled()
            gc.collect()
            gc.disable()  # a collection in the middle of the timed region stalls the launching thread for tens of ms
            barrier()
            _lib.launch_count = 0
            if with_hook:
                _lib.event_hook = hook
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                fn()
            e1.record()
            barrier()
            if was_enabled:
                gc.enable()
            _lib.event_hook = None
            time.sleep(0.15)
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms, _lib.launch_count, cs.summary()

    # (1) device-resident throughput
    ms_dev, launches, clocks = timed(lambda: model.videomae(x_dev), args.steps, with_hook=True)
    attn_ms = [a.elapsed_time(b) for a, b in attn_events]
    attn_avg = sum(attn_ms) / max(len(attn_ms), 1)

    # (2) end to end through the public API (EmbeddingRunner.embed_stream) with HOST buffers: every step copies its
    #     335.5 MB volume from pinned host memory and reads its 62.9 MB embedding back into pinned host memory; the
    #     runner overlaps those copies with the compute of the neighbouring volumes (3 streams, double buffers).
    from smb_vision_b200.inference import EmbeddingRunner

    runner = EmbeddingRunner(model)
    x_hosts = [x_host, x_host.clone().pin_memory()]

    def e2e_run(n):
        tot = 0.0
        for emb in runner.embed_stream(x_hosts[i & 1] for i in range(n)):
            tot += float(emb[0, 0, 0])  # touch the host result
        return tot

    import gc

    e2e_run(args.warmup)
    gc.collect()
    gc.disable()  # (re-enabled after the last timed region) a collection stalls the launching thread for tens of ms
    barrier()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    e2e_run(args.steps)
    e1.record()
    barrier()
    ms_e2e = max(e0.elapsed_time(e1), 0.0)
    wall_e2e = (time.perf_counter() - t0) * 1e3
    ms_e2e = max(ms_e2e, wall_e2e * 0.0 + ms_e2e)
    if world > 1:
        tt = torch.tensor([ms_e2e], device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms_e2e = tt.item()

    # (2b) the same through the raw-volume entry point: int16 HU volumes from pinned host memory (half the PCIe bytes), the
    #      scale / pad / crop / permute tail of the reference dataloader runs on the GPU in front of the encoder
    from smb_vision_b200.data import VolumePreprocessor

    gen_r = torch.Generator().manual_seed(23 + rank)
    raw_inf = [torch.randint(-1100, 1500, (512, 512, 320), generator=gen_r, dtype=torch.int16).pin_memory() for _ in range(2)]
    runner_raw = EmbeddingRunner(model, preprocess=VolumePreprocessor(512, 320, device=dev))

    def e2e_raw_run(n):
        tot = 0.0
        for emb in runner_raw.embed_stream(raw_inf[i & 1] for i in range(n)):
            tot += float(emb[0, 0, 0])
        return tot

    e2e_raw_run(args.warmup)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    e2e_raw_run(args.steps)
    e1.record()
    barrier()
    ms_e2e_raw = e0.elapsed_time(e1)
    if world > 1:
        tt = torch.tensor([ms_e2e_raw], device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms_e2e_raw = tt.item()
    del runner_raw

    # (3) MIM pre-training step (BASELINE configs[2]): forward + loss + backward + bucketed bf16 gradient all-reduce
    #     (NCCL, overlapped with backward) + gradient clipping + AdamW (one fused pass over flat arenas), batch 1 volume per GPU
    mim = None
    if not args.no_mim:
        from smb_vision_b200.modeling import _prep_mask
        from smb_vision_b200.training import DataParallelStep

        np.random.seed(rank)
        mask = torch.from_numpy(OracleMaskGenerator(512, 320, 32, 16, 0.65)()).unsqueeze(0)
        n_mask = int(mask.sum())
        vol_dev = model.videomae._volume(x_dev)
        mp = _prep_mask(mask, dev, n_mask)
        model.train()
        from smb_vision_b200.optim import FusedAdamW

        opt = FusedAdamW(model, lr=5e-5, weight_decay=0.01, max_grad_norm=1.0)  # scripts/training/run_mim.sh:17-21
        dp = DataParallelStep(model, optimizer=opt)
        tsteps = max(args.steps, 3)
        losses = []
        ms_fb, _, _ = timed(lambda: losses.append(dp.step(vol_dev, mp)[0]), tsteps)
        TRAIN_FLOPS = 18.461e12  # 3 x 6.154 TFLOP forward, no recompute
        mim = {"train_step_ms": ms_fb / tsteps, "volumes_per_s": world * tsteps / (ms_fb / 1e3),
               "includes": "forward + norm-pix MSE loss + backward + bf16 gradient all-reduce (world>1) + global-norm clip 1.0 + AdamW (smbv_adamw_step: fp32 master weights, bf16 operand refresh in the same pass)",
               "model_tflops_per_gpu": TRAIN_FLOPS * tsteps / (ms_fb / 1e3) / 1e12,
               "frac_of_sustained_peak": TRAIN_FLOPS * tsteps / (ms_fb / 1e3) / 1e12 / peaks()["tf_sust"],
               "loss_first": float(losses[0]), "loss_last": float(losses[-1]),
               "peak_mem_gb": torch.cuda.max_memory_allocated() / 2**30}

        # (3b) the same step end to end from HOST data through the public API, every step: raw int16 CT volume (pinned host,
        #      167.8 MB) -> H2D (side stream, double-buffered, overlaps the previous step) -> VolumePreprocessor (scale / pad /
        #      crop / permute kernel) -> MaskGenerator.device_batch (fresh mask from the reference RNG stream, index lists on
        #      the GPU) -> DataParallelStep.step -> the loss is read back on the host (one step behind, so the CPU can run ahead)
        from smb_vision_b200.data import MaskGenerator, VolumePreprocessor

        gen = torch.Generator().manual_seed(11 + rank)
        raw_hosts = [torch.randint(-1100, 1500, (512, 512, 320), generator=gen, dtype=torch.int16).pin_memory() for _ in range(2)]
        raw_dev = [torch.empty((512, 512, 320), dtype=torch.int16, device=dev) for _ in range(2)]
        prep, mgen = VolumePreprocessor(512, 320, device=dev), MaskGenerator(512, 320, 32, 16, 0.65)
        s_in = torch.cuda.Stream(dev)
        ev_in, ev_free = [torch.cuda.Event() for _ in range(2)], [torch.cuda.Event() for _ in range(2)]
        loss_host = torch.zeros(2, dtype=torch.float32).pin_memory()
        ev_loss = [torch.cuda.Event() for _ in range(2)]
        seen = []

        def h2d(i):
            k = i & 1
            with torch.cuda.stream(s_in):
                s_in.wait_event(ev_free[k])
                raw_dev[k].copy_(raw_hosts[k], non_blocking=True)
                ev_in[k].record(s_in)

        def e2e_train(n):
            cur = torch.cuda.current_stream(dev)
            h2d(0)
            for i in range(n):
                k = i & 1
                if i + 1 < n:
                    h2d(i + 1)  # next volume's copy runs under this step's compute
                cur.wait_event(ev_in[k])
                vol = prep(raw_dev[k]).view(1, 320, 512, 512)
                ev_free[k].record(cur)
                loss, _ = dp.step(vol, mgen.device_batch(1, dev))
                loss_host[k:k + 1].copy_(loss.reshape(1), non_blocking=True)
                ev_loss[k].record(cur)
                if i > 0:
                    ev_loss[k ^ 1].synchronize()
                    seen.append(float(loss_host[k ^ 1]))
            ev_loss[(n - 1) & 1].synchronize()
            seen.append(float(loss_host[(n - 1) & 1]))

        e2e_train(args.warmup)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        e2e_train(tsteps)
        e1.record()
        barrier()
        ms_te = e0.elapsed_time(e1)
        if world > 1:
            tt = torch.tensor([ms_te], device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ms_te = tt.item()
        mim["e2e"] = {"value": world * tsteps / (ms_te / 1e3), "unit": UNIT, "ms_per_step": ms_te / tsteps,
                      "h2d_bytes_per_step": raw_hosts[0].numel() * 2 + 2560, "d2h_bytes_per_step": 4,
                      "api": "VolumePreprocessor + MaskGenerator.device_batch + DataParallelStep.step(FusedAdamW): raw int16 volume from pinned "
                             "host memory each step, fresh mask each step, loss read back on the host",
                      "loss_last": seen[-1]}
        model.eval()

    # (4) SURVEY.md §8f rank 4: V-JEPA2-3D ViT-L encoder forward on the same volume (embedding extraction / the momentum
    #     target encoder's pass of every V-JEPA step, src/run_vjepa.py:126-135).  Secondary number: it must not take the
    #     headline line down with it, so a failure is reported inside the block.
    vjepa = None
    if not args.no_vjepa:
        try:
            from transformers import VJEPA2Config

            from smb_vision_b200.vjepa import B200VJEPA2Model

            vc = VJEPA2Config(patch_size=16, crop_size=512, frames_per_clip=320, tubelet_size=16, in_chans=1)  # src/run_vjepa.py:220-232
            torch.manual_seed(1)
            with torch.device(dev):
                vmodel = B200VJEPA2Model(vc, with_predictor=False).eval()
            vsteps = max(min(args.steps, 5), 3)
            ms_vj, n_launch, _ = timed(lambda: vmodel.get_vision_features(x_dev), vsteps)
            VJ_FLOPS = 2 * 20480 * 4096 * 1024 + 24 * (2 * 20480 * 1024 * 12 * 1024 + 4 * 20480 * 20480 * 1024)
            vjepa = {"workload": "V-JEPA2-3D ViT-L (1024/16 heads/24 layers) encoder forward, 512x512x320 = 20480 tokens, batch 1/GPU, bf16, random init",
                     "volumes_per_s": world * vsteps / (ms_vj / 1e3), "ms_per_volume": ms_vj / vsteps, "gpu_launches": n_launch,
                     "model_tflops_per_gpu": VJ_FLOPS * vsteps / (ms_vj / 1e3) / 1e12,
                     "frac_of_sustained_peak": VJ_FLOPS * vsteps / (ms_vj / 1e3) / 1e12 / peaks()["tf_sust"]}
            del vmodel
        except Exception as e:  # noqa: BLE001
            vjepa = {"error": f"{type(e).__name__}: {e}"}

    gc.enable()
    pk = peaks()
    vps = world * args.steps / (ms_dev / 1e3)
    vps_e2e = world * args.steps / (ms_e2e / 1e3)
    ach = ATTN_FLOPS_PER_LAUNCH / (attn_avg / 1e3) / 1e12 if attn_avg > 0 else 0.0
    line = {
        "metric": METRIC, "value": vps, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": {"workload": WORKLOAD, "parallelism": f"volume-sharded x{world}, no collective", "l2": "inputs larger than L2 (335.5 MB volume, 63 MB activations per op)",
                   "tflops_per_volume": EMBED_FLOPS / 1e12},
        "model_tflops": EMBED_FLOPS * vps / world / 1e12,
        "model_frac_of_sustained_peak": EMBED_FLOPS * vps / world / 1e12 / pk["tf_sust"],
        "e2e": {"value": vps_e2e, "unit": UNIT, "h2d_bytes_per_step": x_host.numel() * 4, "d2h_bytes_per_step": emb_host.numel() * 4,
                "ms_per_step": ms_e2e / args.steps, "api": "smb_vision_b200.inference.EmbeddingRunner.embed_stream (pinned host in, pinned host out, copies overlapped with compute)"},
        "e2e_raw_int16": {"value": world * args.steps / (ms_e2e_raw / 1e3), "unit": UNIT, "h2d_bytes_per_step": raw_inf[0].numel() * 2,
                          "d2h_bytes_per_step": emb_host.numel() * 4, "ms_per_step": ms_e2e_raw / args.steps,
                          "api": "EmbeddingRunner(model, preprocess=VolumePreprocessor(512, 320)).embed_stream: raw int16 HU volume in, fp32 embedding out"},
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": {"kernel": "flash_attn_fwd2_kernel (H=12, N=20480, d=64)", "bound": "tensor", "achieved": ach, "peak": pk["tf_sust"],
                     "unit": "TFLOP/s", "frac": ach / pk["tf_sust"], "traffic": ATTN_DRAM_BYTES_PER_LAUNCH,
                     "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture of this kernel, per launch "
                                       "(profiles/r01_ncu_full.md); algorithmic Q,K,V,O bytes = 125.8 MB",
                     "peak_source": pk["src"] + " sustained bf16",
                     "launch_ms": attn_avg, "launches_timed": len(attn_ms), "share_of_step": attn_avg * LAYERS / (ms_dev / args.steps)},
    }
    if mim:
        line["mim"] = mim
    if vjepa:
        line["vjepa_encoder"] = vjepa
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        t, cores, sample = cpu_embed_sample(1)
        line["cpu_baseline"] = {"value": 1.0 / t, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
