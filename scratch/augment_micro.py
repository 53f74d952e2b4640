"""GPU input pipeline throughput: 512 decoded 256x256 uint8 images -> 128x128 fp32 model input, both loaders.
usage: python scratch/augment_micro.py [batch]"""
import sys
sys.path.insert(0, ".")
import torch
from locate_b200.augment import GpuAugment

b = int(sys.argv[1]) if len(sys.argv) > 1 else 512
img = torch.randint(0, 256, (b, 256, 256, 3), dtype=torch.uint8, device="cuda")
aug = GpuAugment(128, seed=0)
for augmented in (False, True):
    prm = aug.draw(b, 256, 256, augmented).cuda()
    for _ in range(3):
        aug(img, augmented, params=prm)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        aug(img, augmented, params=prm)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    byts = img.numel() * 2 + b * 3 * 128 * 128 * 4          # source read twice (mean pass + apply) + output
    print(f"{'augmented' if augmented else 'base     '} loader: {ms:.3f} ms per batch of {b} = {b / ms * 1e3:,.0f} images/s, {byts / ms / 1e6:.0f} GB/s")
