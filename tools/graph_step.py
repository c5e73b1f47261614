"""Does a CUDA graph of the whole MIM training step beat the 475 individual launches?  (launch-gap measurement)
usage: python tools/graph_step.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from __graft_entry__ import hf_config
from oracle import videomae_oracle as vo
from oracle.mim_mask import OracleMaskGenerator
from smb_vision_b200.modeling import B200VideoMAEForPreTraining, _prep_mask
from smb_vision_b200.training import DataParallelStep
from smb_vision_b200.optim import FusedAdamW

dev = torch.device("cuda", 0)
c = vo.OracleConfig()
torch.manual_seed(1234)
model = B200VideoMAEForPreTraining(hf_config({k: getattr(c, k) for k in c.__dataclass_fields__})).to(dev).train()
vol = model.videomae._volume(vo.synthetic_volume(c, 1, 7).to(dev))
np.random.seed(0)
mask = torch.from_numpy(np.stack([OracleMaskGenerator(512, 320, 32, 16, 0.65)()]))
mp = _prep_mask(mask, dev, int(mask[0].sum()))
dp = DataParallelStep(model, optimizer=FusedAdamW(model, lr=5e-5, weight_decay=0.01, max_grad_norm=1.0))

def timeit(fn, n=10):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

for _ in range(3): loss, _ = dp.step(vol, mp)
print("eager step", round(timeit(lambda: dp.step(vol, mp)), 3), "ms", flush=True)
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for _ in range(2): dp.step(vol, mp)
torch.cuda.current_stream().wait_stream(s)
g = torch.cuda.CUDAGraph()
try:
    with torch.cuda.graph(g):
        out = dp.step(vol, mp)
    print("captured", flush=True)
    for _ in range(3): g.replay()
    print("graph step", round(timeit(g.replay), 3), "ms; loss", float(out[0]), flush=True)
except Exception as e:
    print("capture failed:", repr(e)[:600], flush=True)
