"""SEPARABLE = True (reference libs/config.py:53, SURVEY.md 'next' row N3): the depthwise and grouped full-extent
convolution kernels of csrc/depthwise.cu through the C ABI against torch's grouped convolutions in fp32 on the same
operands (bf16 storage: both sides read the same bf16-rounded tensors, so only accumulation order and the output
rounding differ)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from locate_b200._lib import call, ptr  # noqa: E402

DEV = "cuda:0"
F32, BF16 = 0, 1


def cl(t, dtype):
    """NCHW values -> channels-last storage [B][H][W][C] in `dtype`"""
    return t.permute(0, 2, 3, 1).contiguous().to(dtype)


def nchw(t):
    return t.float().permute(0, 3, 1, 2).contiguous()


def tol(dt):
    return dict(rtol=1e-4, atol=1e-4) if dt == F32 else dict(rtol=1.2e-2, atol=2e-2)


DW = [  # b, c, h, w, k, stride, pad, transposed
    (2, 8, 9, 7, 3, 1, 1, False),
    (3, 48, 16, 16, 3, 1, 1, False),
    (2, 96, 16, 16, 5, 2, 2, False),
    (2, 24, 13, 11, 5, 2, 2, False),
    (2, 96, 8, 8, 4, 2, 1, True),
    (1, 40, 5, 6, 4, 2, 1, True),
    (2, 5, 6, 6, 3, 1, 1, False),          # channel count with no vector width
    (1, 768, 4, 4, 3, 1, 1, False),
]


@pytest.mark.parametrize("dt", [F32, BF16])
@pytest.mark.parametrize("case", DW)
def test_depthwise_forward_dgrad_wgrad(case, dt):
    b, c, h, w, k, s, p, tr = case
    gen = torch.Generator().manual_seed(c * 131 + k)
    tdt = torch.float32 if dt == F32 else torch.bfloat16
    x = cl(torch.randn(b, c, h, w, generator=gen), tdt).to(DEV)
    wt = torch.randn(c, 1, k, k, generator=gen).to(DEV)
    alpha = torch.tensor([0.37], device=DEV)
    xr = nchw(x)
    if tr:
        ref = F.conv_transpose2d(xr, wt, stride=s, padding=p, groups=c) * alpha
    else:
        ref = F.conv2d(xr, wt, stride=s, padding=p, groups=c) * alpha
    oh, ow = ref.shape[2:]
    out = torch.empty(b, oh, ow, c, dtype=tdt, device=DEV)
    call("lb_dw_conv", ptr(x), ptr(wt), ptr(alpha), ptr(out), b, h, w, oh, ow, c, k, k, s, p, 1 if tr else 0, dt)
    torch.testing.assert_close(nchw(out), ref, **tol(dt))

    # input gradient = the other mode on the same weights
    g = cl(torch.randn(b, c, oh, ow, generator=gen), tdt).to(DEV)
    gr = nchw(g)
    if tr:
        ref_dx = F.conv2d(gr, wt, stride=s, padding=p, groups=c) * alpha
    else:
        ref_dx = torch.nn.grad.conv2d_input(xr.shape, wt, gr, stride=s, padding=p, groups=c) * alpha
    dx = torch.empty_like(x)
    call("lb_dw_conv", ptr(g), ptr(wt), ptr(alpha), ptr(dx), b, oh, ow, h, w, c, k, k, s, p, 0 if tr else 1, dt)
    torch.testing.assert_close(nchw(dx), ref_dx, **tol(dt))

    # weight gradient (accumulating)
    if tr:
        ref_dw = torch.nn.grad.conv2d_weight(gr, wt.shape, xr, stride=s, padding=p, groups=c)
        args = (ptr(g), ptr(x), None, b, oh, ow, h, w)
    else:
        ref_dw = torch.nn.grad.conv2d_weight(xr, wt.shape, gr, stride=s, padding=p, groups=c)
        args = (ptr(x), ptr(g), None, b, h, w, oh, ow)
    dw = torch.full_like(wt, 0.5)
    call("lb_dw_wgrad", args[0], args[1], ptr(dw), *args[3:], c, k, k, s, p, dt)
    scale = ref_dw.abs().max().item()
    torch.testing.assert_close(dw - 0.5, ref_dw, rtol=1e-3, atol=2e-5 * max(scale, 1.0) * (1 if dt == F32 else 50))


GF = [  # b, features, r, size
    (3, 8, 2, 4),
    (2, 96, 2, 8),
    (4, 48, 2, 16),
    (2, 768, 2, 2),
    (2, 12, 4, 3),
    (130, 96, 2, 4),
]


@pytest.mark.parametrize("dt", [F32, BF16])
@pytest.mark.parametrize("case", GF)
def test_grouped_full_extent_conv(case, dt):
    b, f, r, size = case
    gen = torch.Generator().manual_seed(f * 7 + size)
    tdt = torch.float32 if dt == F32 else torch.bfloat16
    x = cl(torch.randn(b, f, size, size, generator=gen), tdt).to(DEV)
    wt = torch.randn(f // r, r, size, size, generator=gen).to(DEV)
    alpha = torch.tensor([1.7], device=DEV)
    xr = nchw(x)
    ref = F.conv2d(xr, wt, groups=f // r) * alpha                      # [B, F/r, 1, 1]
    out = torch.empty(b, f // r, dtype=tdt, device=DEV)
    call("lb_gfull_fwd", ptr(x), ptr(wt), ptr(alpha), ptr(out), b, size * size, f, r, dt)
    torch.testing.assert_close(out.float(), ref.reshape(b, f // r), **tol(dt))

    g = torch.randn(b, f // r, generator=gen).to(tdt).to(DEV)
    gr = g.float().reshape(b, f // r, 1, 1)
    ref_dx = torch.nn.grad.conv2d_input(xr.shape, wt, gr, groups=f // r) * alpha
    dx = torch.empty_like(x)
    call("lb_gfull_dgrad", ptr(g), ptr(wt), ptr(alpha), ptr(dx), b, size * size, f, r, dt)
    torch.testing.assert_close(nchw(dx), ref_dx, **tol(dt))

    ref_dw = torch.nn.grad.conv2d_weight(xr, wt.shape, gr, groups=f // r)
    dw = torch.full_like(wt, -0.25)
    call("lb_gfull_wgrad", ptr(x), ptr(g), ptr(dw), b, size * size, f, r, dt)
    torch.testing.assert_close(dw + 0.25, ref_dw, rtol=1e-3, atol=1e-4 * max(ref_dw.abs().max().item(), 1.0))
    # fixed summation order over the batch: bit-identical on a second run
    dw2 = torch.full_like(wt, -0.25)
    call("lb_gfull_wgrad", ptr(x), ptr(g), ptr(dw2), b, size * size, f, r, dt)
    assert torch.equal(dw, dw2)
