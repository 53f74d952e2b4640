"""CPU restatement of the reference's input transforms for ONE sample with EXPLICIT random draws -- test infrastructure
only (never imported by product code).

Follows libs/utils.py:92-113: Resize(2S) is assumed done; then (augmented only) RandomHorizontalFlip and
ColorJitter(brightness, contrast, saturation) in a drawn order, RandomResizedCrop -> resize to S with antialiased
bilinear interpolation, ToTensor, Normalize((0.5,)*3, (0.5,)*3).  The arithmetic is torchvision's float-tensor path
(torchvision/transforms/_functional_tensor.py: _blend, rgb_to_grayscale, adjust_*), which tests/test_augment_oracle.py
pins against torchvision itself.  parity note: PIL's path (what ImageFolder feeds the reference) rounds to uint8 after
every operation; the restatement and the kernels do not (difference <= 1.5/255 per operation)."""
import torch
import torch.nn.functional as F


def gray(img):                       # img [3, H, W] in [0, 1]
    return (0.2989 * img[0] + 0.587 * img[1] + 0.114 * img[2]).unsqueeze(0)


def blend(a, b, f):
    return (f * a + (1.0 - f) * b).clamp(0.0, 1.0)


def transform(img_u8, params, size):
    """img_u8: uint8 [H, W, 3]; params: the 12 floats of csrc/augment.cu; returns fp32 [3, size, size]."""
    img = img_u8.permute(2, 0, 1).to(torch.float32) / 255.0
    top, left, ch, cw, flip = (int(params[i]) for i in range(5))
    if flip:
        img = img.flip(-1)
    for code in params[8:11].tolist():
        code = int(code)
        if code == 0:
            img = blend(img, torch.zeros_like(img), float(params[5]))
        elif code == 1:
            img = blend(img, gray(img).mean(), float(params[6]))
        elif code == 2:
            img = blend(img, gray(img), float(params[7]))
    crop = img[:, top:top + ch, left:left + cw]
    out = F.interpolate(crop.unsqueeze(0), size=(size, size), mode="bilinear", align_corners=False, antialias=True)[0]
    return (out - 0.5) / 0.5
