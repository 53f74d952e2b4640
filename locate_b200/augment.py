"""GPU-side input pipeline (SURVEY.md "next" row N4): the reference's two loaders (libs/utils.py:92-113,
main.py:122-129) as one batched kernel call.

    base      = Resize(2S) -> RandomResizedCrop(S, (min_crop_part, 1), (1, 1)) -> ToTensor -> Normalize(0.5, 0.5)
    augmented = Resize(2S) -> RandomHorizontalFlip -> ColorJitter(j, j, j) -> RandomResizedCrop(...) -> ToTensor -> Normalize

The decode and the fixed Resize(2S) stay where the files are (host); from there on a batch of uint8 images lives in HBM
and `GpuAugment` produces the model's input tensor directly.  Random draws follow torchvision's `get_params` of each
transform (same distributions, drawn on the host with a torch.Generator); the arithmetic is torchvision's float-tensor
path (no rounding to uint8 between operations, which PIL's path has)."""
import math

import torch

from ._lib import call, ptr

N_PARAMS = 12


class GpuAugment:
    def __init__(self, image_size, jitter=0.2, min_crop_part=0.75, seed=None):
        self.size, self.jitter, self.min_crop = int(image_size), float(jitter), float(min_crop_part)
        self.gen = torch.Generator()
        if seed is not None:
            self.gen.manual_seed(seed)

    # -- torchvision.transforms.RandomResizedCrop.get_params with scale=(min_crop_part, 1), ratio=(1, 1)
    def _crop(self, height, width):
        area = height * width
        for _ in range(10):
            target = area * float(torch.empty(1).uniform_(self.min_crop, 1.0, generator=self.gen))
            aspect = math.exp(float(torch.empty(1).uniform_(0.0, 0.0, generator=self.gen)))    # log-uniform over (1, 1)
            w = int(round(math.sqrt(target * aspect)))
            h = int(round(math.sqrt(target / aspect)))
            if 0 < w <= width and 0 < h <= height:
                i = int(torch.randint(0, height - h + 1, (1,), generator=self.gen))
                j = int(torch.randint(0, width - w + 1, (1,), generator=self.gen))
                return i, j, h, w
        side = min(height, width)                           # fallback: central crop (ratio bounds are 1)
        return (height - side) // 2, (width - side) // 2, side, side

    def draw(self, batch, height, width, augmented):
        """[batch, 12] fp32 host tensor of per-sample parameters (layout: csrc/augment.cu)."""
        p = torch.zeros((batch, N_PARAMS), dtype=torch.float32)
        p[:, 5:8] = 1.0
        p[:, 8:11] = -1.0
        for b in range(batch):
            if augmented:
                p[b, 4] = float(torch.rand(1, generator=self.gen) < 0.5)                          # RandomHorizontalFlip(p=0.5)
                order = torch.randperm(4, generator=self.gen).tolist()                              # ColorJitter.get_params
                lo, hi = max(0.0, 1.0 - self.jitter), 1.0 + self.jitter
                p[b, 5:8] = torch.empty(3).uniform_(lo, hi, generator=self.gen)
                p[b, 8:11] = torch.tensor([o for o in order if o != 3], dtype=torch.float32)        # index 3 = hue: not used (hue=None)
            p[b, 0:4] = torch.tensor(self._crop(height, width), dtype=torch.float32)
        return p

    def __call__(self, images_u8, augmented, params=None):
        """images_u8: uint8 CUDA tensor [B, H, W, 3]; returns logical [B, 3, S, S] fp32 (channels-last storage) in [-1, 1]."""
        if images_u8.dtype != torch.uint8 or images_u8.dim() != 4 or images_u8.shape[-1] != 3 or not images_u8.is_cuda:
            raise ValueError("expected a uint8 CUDA tensor [B, H, W, 3]")
        images_u8 = images_u8.contiguous()
        b, h, w, _ = images_u8.shape
        if params is None:
            params = self.draw(b, h, w, augmented)
        params = params.to(images_u8.device, non_blocking=True).contiguous()
        out = torch.empty((b, 3, self.size, self.size), dtype=torch.float32, device=images_u8.device).contiguous(
            memory_format=torch.channels_last)
        work = torch.empty(b, dtype=torch.float32, device=images_u8.device)
        call("lb_augment", ptr(images_u8), ptr(params), ptr(work), ptr(out), b, h, w, self.size)
        return out
