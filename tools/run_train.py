"""Run a few full-size MIM training steps (for ncu launch lists).  usage: python tools/run_train.py [steps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from __graft_entry__ import hf_config
from oracle import videomae_oracle as vo
from oracle.mim_mask import OracleMaskGenerator
from smb_vision_b200.modeling import B200VideoMAEForPreTraining, _prep_mask
from smb_vision_b200.training import DataParallelStep

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1  # volumes per GPU per step (scripts/training/run_mim.sh uses 4)
dev = torch.device("cuda", 0)
c = vo.OracleConfig()
torch.manual_seed(1234)
model = B200VideoMAEForPreTraining(hf_config({k: getattr(c, k) for k in c.__dataclass_fields__})).to(dev).train()
vol = model.videomae._volume(vo.synthetic_volume(c, B, 7).to(dev))
np.random.seed(0)
_g = OracleMaskGenerator(512, 320, 32, 16, 0.65)
mask = torch.from_numpy(np.stack([_g() for _ in range(B)]))
mp = _prep_mask(mask, dev, int(mask[0].sum()))
from smb_vision_b200.optim import FusedAdamW
dp = DataParallelStep(model, optimizer=FusedAdamW(model, lr=5e-5, weight_decay=0.01, max_grad_norm=1.0))
for _ in range(2):
    dp.step(vol, mp)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    loss, _ = dp.step(vol, mp)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
print(f"train step (batch {B}): {ms:.3f} ms = {B / ms * 1e3:.2f} volumes/s  loss {float(loss):.6f}  peak mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GB")
