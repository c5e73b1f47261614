"""A/B of the second-stream weight-gradient queue (training.SideQueue): SMBV_WGRAD_STREAM=1 vs 0, eager and CUDA-graph step,
interleaved in one process on one box; also checks that both settings leave the same gradients (full-size MIM step).
usage: python tools/side_stream_ab.py [rounds]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from __graft_entry__ import hf_config
from oracle import videomae_oracle as vo
from smb_vision_b200.data import MaskGenerator
from smb_vision_b200.modeling import B200VideoMAEForPreTraining
from smb_vision_b200.training import DataParallelStep, GradArena, mim_backward, mim_forward_train
from smb_vision_b200.optim import FusedAdamW

rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 3
dev = torch.device("cuda", 0)
c = vo.OracleConfig()
torch.manual_seed(1234)
model = B200VideoMAEForPreTraining(hf_config({k: getattr(c, k) for k in c.__dataclass_fields__})).to(dev).train()
vol = model.videomae._volume(vo.synthetic_volume(c, 1, 7).to(dev))
np.random.seed(0)
mp = MaskGenerator(512, 320, 32, 16, 0.65).device_batch(1, dev)


def timeit(fn, n=10):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


# gradients: side stream on (5 times) vs off
with torch.no_grad():
    loss, logits, dlogits, S = mim_forward_train(model, vol, mp)
    os.environ["SMBV_WGRAD_STREAM"] = "0"
    a0 = GradArena(model, dev); mim_backward(model, S, dlogits, a0)
    b0 = GradArena(model, dev); mim_backward(model, S, dlogits, b0)
    torch.cuda.synchronize()
    print("off vs off: max |d|", float((a0.flat - b0.flat).abs().max()), "rel", float(torch.linalg.norm(a0.flat - b0.flat) / torch.linalg.norm(a0.flat)))
    os.environ["SMBV_WGRAD_STREAM"] = "1"
    for i in range(5):
        a1 = GradArena(model, dev); mim_backward(model, S, dlogits, a1)
        torch.cuda.synchronize()
        print("on vs off: max |d|", float((a1.flat - a0.flat).abs().max()), "rel", float(torch.linalg.norm(a1.flat - a0.flat) / torch.linalg.norm(a0.flat)), flush=True)
    del S, a0, b0, a1

steps = {}
for mode in ("0", "1"):
    os.environ["SMBV_WGRAD_STREAM"] = mode
    for graph in (False, True):
        dp = DataParallelStep(model, optimizer=FusedAdamW(model, lr=5e-5, weight_decay=0.01, max_grad_norm=1.0), cuda_graph=graph)
        for _ in range(3): dp.step(vol, mp)
        steps[(mode, graph)] = dp
for r in range(rounds):
    for (mode, graph), dp in steps.items():
        os.environ["SMBV_WGRAD_STREAM"] = mode
        t = timeit(lambda: dp.step(vol, mp))
        print(f"round {r} side_stream={mode} graph={graph}: {t:.3f} ms/step", flush=True)
