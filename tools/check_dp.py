"""Multi-GPU check of the data-parallel MIM step (run under torchrun, one process per GPU):
  * DataParallelStep + FusedAdamW keep the replicas bit-identical (same all-reduced gradients -> same update);
  * N ranks x batch 1 == one process with batch N (mean of per-rank mean losses = batch mean loss), up to the bf16 wire.
usage: python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/check_dp.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import __graft_entry__ as ge
from oracle import videomae_oracle as vo
from oracle.mim_mask import OracleMaskGenerator
from smb_vision_b200.modeling import B200VideoMAEForPreTraining, _prep_mask
from smb_vision_b200.optim import FusedAdamW
from smb_vision_b200.training import DataParallelStep

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
cfg = vo.OracleConfig(**ge.SMALL64)
sd = vo.synthetic_state_dict(cfg, 1234)
xs = vo.synthetic_volume(cfg, world, 7)
np.random.seed(0)
g = OracleMaskGenerator(96, 96, 32, 16, 0.65)
masks = torch.from_numpy(np.stack([g() for _ in range(world)]))


def run(model_batch, masks_b, group):
    m = B200VideoMAEForPreTraining(ge.hf_config(ge.SMALL64)).to(dev)
    opt = FusedAdamW(m, lr=1e-3, weight_decay=0.01, max_grad_norm=1.0)
    m.load_state_dict(sd, strict=True)
    dp = DataParallelStep(m, optimizer=opt, process_group=group)
    vol = m.videomae._volume(model_batch.to(dev))
    mp = _prep_mask(masks_b, dev, None)
    losses = [dp.step(vol, mp)[0].item() for _ in range(3)]
    return m, losses, opt


# (a) data parallel: rank r trains on sample r
m_dp, losses_dp, opt = run(xs[rank:rank + 1], masks[rank:rank + 1], None)
flat = opt.params.flat
ref = flat.clone()
dist.broadcast(ref, src=0)
same = bool(torch.equal(ref, flat))
allsame = torch.tensor([1.0 if same else 0.0], device=dev)
dist.all_reduce(allsame, op=dist.ReduceOp.MIN)
mean_loss = torch.tensor(losses_dp, device=dev)
dist.all_reduce(mean_loss)
mean_loss /= world
# (b) one process, batch = world (no communication): every rank computes it locally with world-size-1 semantics
dist.barrier()
m1 = B200VideoMAEForPreTraining(ge.hf_config(ge.SMALL64)).to(dev)
opt1 = FusedAdamW(m1, lr=1e-3, weight_decay=0.01, max_grad_norm=1.0)
m1.load_state_dict(sd, strict=True)
import smb_vision_b200.training as tr
dp1 = tr.DataParallelStep(m1, optimizer=opt1)
dp1.reducer.world = 1  # batch-N reference: no all-reduce
vol = m1.videomae._volume(xs.to(dev))
mp = _prep_mask(masks, dev, None)
losses_1 = [dp1.step(vol, mp)[0].item() for _ in range(3)]
num = (opt1.params.flat - flat).norm().item()
den = opt1.params.flat.norm().item()
if rank == 0:
    print(f"world {world}: replicas identical after 3 steps: {bool(allsame.item())}")
    print(f"mean DP losses {[round(v, 6) for v in mean_loss.tolist()]} vs batch-{world} losses {[round(v, 6) for v in losses_1]}")
    print(f"params DP vs batch-{world}: frob-rel {num / den:.3e}")
    ok = bool(allsame.item()) and all(abs(a - b) / b < 2e-3 for a, b in zip(mean_loss.tolist(), losses_1)) and num / den < 2e-3
    print("CHECK_DP", "PASS" if ok else "FAIL")
dist.barrier()
dist.destroy_process_group()
