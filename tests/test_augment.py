"""SURVEY "next" row N4: the GPU input pipeline.  CPU: the restatement (oracle/augment_oracle.py) against torchvision's own
functional transforms (the reference's dependency) and the parameter draws against the transforms' documented ranges;
GPU: lb_augment through the C ABI against the restatement."""
import pytest
import torch

from oracle import augment_oracle as AO


def _images(b, h, w, seed=0):
    g = torch.Generator().manual_seed(seed)
    base = torch.rand((b, h, w, 3), generator=g)
    yy = torch.linspace(0, 1, h).view(1, h, 1, 1)
    return ((0.6 * base + 0.4 * yy) * 255).round().to(torch.uint8)


def test_restatement_matches_torchvision_functional():
    tvf = pytest.importorskip("torchvision.transforms.v2.functional")
    img = _images(1, 40, 48)[0]
    prm = torch.tensor([3, 5, 30, 30, 1, 1.13, 0.84, 1.19, 2, 0, 1, 0], dtype=torch.float32)
    got = AO.transform(img, prm, 16)
    x = img.permute(2, 0, 1).float() / 255
    x = tvf.horizontal_flip(x)
    x = tvf.adjust_saturation(x, 1.19)
    x = tvf.adjust_brightness(x, 1.13)
    x = tvf.adjust_contrast(x, 0.84)
    x = tvf.resized_crop(x, 3, 5, 30, 30, [16, 16], antialias=True)
    x = tvf.normalize(x, [0.5] * 3, [0.5] * 3)
    assert torch.allclose(got, x, atol=2e-6), (got - x).abs().max()


def test_parameter_draws_follow_the_transforms():
    from locate_b200.augment import GpuAugment
    aug = GpuAugment(64, jitter=0.2, min_crop_part=0.75, seed=1)
    p = aug.draw(256, 128, 160, augmented=True)
    area = p[:, 2] * p[:, 3] / (128 * 160)
    assert torch.equal(p[:, 2], p[:, 3])                                           # ratio (1, 1): square crops
    assert area.min() > 0.74 and area.max() <= 1.0 + 1e-6 and (p[:, 2] <= 128).all()
    assert ((p[:, 0] >= 0) & (p[:, 0] + p[:, 2] <= 128) & (p[:, 1] >= 0) & (p[:, 1] + p[:, 3] <= 160)).all()
    assert 0.3 < p[:, 4].mean() < 0.7
    assert (p[:, 5:8] >= 0.8).all() and (p[:, 5:8] <= 1.2).all()
    assert all(sorted(r.tolist()) == [0.0, 1.0, 2.0] for r in p[:, 8:11])
    q = aug.draw(8, 128, 160, augmented=False)                                      # base loader: crop only
    assert (q[:, 4] == 0).all() and (q[:, 5:8] == 1).all() and (q[:, 8:11] == -1).all()


@pytest.mark.gpu
@pytest.mark.parametrize("augmented", [False, True])
@pytest.mark.parametrize("h,w,size", [(64, 64, 32), (70, 90, 32), (256, 256, 128), (40, 40, 32)])
def test_gpu_pipeline_matches_restatement(h, w, size, augmented):
    from locate_b200.augment import GpuAugment
    b = 5
    img = _images(b, h, w, seed=h + w)
    aug = GpuAugment(size, seed=7)
    prm = aug.draw(b, h, w, augmented)
    out = aug(img.cuda(), augmented, params=prm)
    assert out.shape == (b, 3, size, size) and out.is_contiguous(memory_format=torch.channels_last)
    for i in range(b):
        want = AO.transform(img[i], prm[i], size)
        assert torch.allclose(out[i].cpu(), want, atol=3e-5), (i, (out[i].cpu() - want).abs().max().item())


@pytest.mark.gpu
def test_gpu_pipeline_feeds_the_discriminator():
    import locate_b200 as L
    from locate_b200.augment import GpuAugment
    L.configure(IMAGE_SIZE=32, BASE_FEATURE_FACTOR=4)
    torch.manual_seed(0)
    dis, _ = L.get_model(L.Discriminator(), L.CFG.DLR, "cuda:0")
    x = GpuAugment(32, seed=3)(_images(4, 64, 64).cuda(), augmented=True)
    with torch.no_grad():
        y = dis(x)
    assert y.shape[0] == 4 and torch.isfinite(y.float()).all()
