"""Classification fine-tuning step at BASELINE.json configs[3]: smb-vision-base encoder, 224x224x160 (1960 tokens), batch 4 per GPU,
2 additional features (age / sex), 2 labels; forward + loss + backward + clip + AdamW.  usage: python tools/run_cls.py [steps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from transformers import VideoMAEConfig
from smb_vision_b200.modeling import B200VideoMAEForVideoClassification
from smb_vision_b200.optim import FusedAdamW
from smb_vision_b200.training import DataParallelStep

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
dev = torch.device("cuda", 0)
c = VideoMAEConfig()
c.update(dict(image_size=224, patch_size=16, num_channels=1, num_frames=160, tubelet_size=16, num_labels=2, additional_features_size=2,
              problem_type="single_label_classification"))  # src/run_classification.py:452-478
torch.manual_seed(0)
model = B200VideoMAEForVideoClassification(c).to(dev).train()
B = 4
x = torch.rand(B, 160, 1, 224, 224, device=dev)
feats = torch.randn(B, 2, device=dev)
labels = torch.randint(0, 2, (B,), device=dev)
opt = FusedAdamW(model, lr=5e-5, weight_decay=0.01, max_grad_norm=1.0)
dp = DataParallelStep(model, optimizer=opt)
vol = model.videomae._volume(x)
for _ in range(3):
    loss, _ = dp.step(vol, feats, labels)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    loss, _ = dp.step(vol, feats, labels)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
N, d, L, mlp = 1960, 768, 12, 3072
flops = 3 * B * (2 * N * 4096 * d + L * (2 * N * d * (3 * d + d + 2 * mlp) + 4 * N * N * 64 * 12))
print(f"classification step (B={B}, N={N}): {ms:.3f} ms = {B / ms * 1e3:.1f} volumes/s, {flops / ms / 1e9:.0f} TFLOP/s, loss {float(loss):.4f}")
with torch.no_grad():
    model.eval()
    for _ in range(3):
        out = model(x, additional_features=feats)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(steps):
        out = model(x, additional_features=feats)
    e1.record()
    torch.cuda.synchronize()
print(f"classification inference (B={B}): {e0.elapsed_time(e1) / steps:.3f} ms")
