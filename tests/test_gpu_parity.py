"""GPU parity: the CUDA path (through the C ABI) vs (a) the golden outputs of the unmodified reference
and (b) the CPU oracle on seeded inputs.  fp32 kernels vs fp32 reference: rtol 1e-4;
accumulation order differs (atomics, tiling), so atol scales with the tensor's magnitude."""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu

import locate_b200 as L                      # noqa: E402
from locate_b200 import layers, ops          # noqa: E402
from oracle import locate_oracle as O        # noqa: E402

DEV = "cuda:0"


def close(a, b, rtol=1e-4, atol_frac=2e-5, what=""):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    assert a.shape == b.shape, f"{what}: shape {tuple(a.shape)} vs {tuple(b.shape)}"
    atol = atol_frac * max(b.abs().max().item(), 1e-30)
    bad = (a - b).abs() > atol + rtol * b.abs()
    assert not bad.any(), (f"{what}: {int(bad.sum())}/{bad.numel()} off, max abs err "
                           f"{(a - b).abs().max().item():.3e}, ref max {b.abs().max().item():.3e}")


def dev(t):
    return t.to(DEV)


@pytest.fixture(autouse=True)
def _cfg():
    L.config.reset()
    L.configure(PRECISION="fp32")       # this file pins the fp32 path; tests/test_gpu_bf16.py covers the tensor-core path
    yield
    L.config.reset()


# ------------------------------------------------------------------------------------------ primitives
def test_roottanh_golden(golden):
    r = golden("primitives.pt")["roottanh"]
    x = dev(r["x"]).requires_grad_(True)
    y = layers.nonlinear_function(x)
    y.backward(dev(r["g"]))
    close(y, r["y"], what="y")
    close(x.grad, r["dx"], what="dx")


def test_roottanh_large_random():
    x = torch.randn(3, 96, 33, 17, generator=torch.Generator().manual_seed(1)) * 4
    g = torch.randn(x.shape, generator=torch.Generator().manual_seed(2))
    xo = x.clone().requires_grad_(True)
    yo = O.roottanh(xo)
    yo.backward(g)
    xd = dev(x).requires_grad_(True)
    yd = layers.nonlinear_function(xd)
    yd.backward(dev(g))
    close(yd, yo, what="y")
    close(xd.grad, xo.grad, what="dx")


@pytest.mark.parametrize("tag", ["norm_plain", "norm_styled"])
def test_norm_golden(golden, tag):
    r = golden("primitives.pt")[tag]
    m = layers.InPlaceNorm(6).to(DEV)
    m.weight.data.copy_(r["weight"])
    m.bias.data.copy_(r["bias"])
    x = dev(r["x"]).requires_grad_(True)
    scale = dev(r["scale"]).requires_grad_(True) if tag == "norm_styled" else None
    y = m(x, scale)
    y.backward(dev(r["g"]))
    close(y, r["y"], what="y")
    close(x.grad, r["dx"], what="dx")
    close(m.bias.grad, r["dbias"], what="dbias")
    if scale is None:
        close(m.weight.grad, r["dweight"], what="dweight")
    else:
        close(scale.grad, r["dscale"], what="dscale")
        assert m.weight.grad is None


@pytest.mark.parametrize("shape,styled", [((5, 96, 16, 16), True), ((2, 30, 7, 9), False), ((1, 3, 64, 64), False)])
def test_norm_vs_oracle(shape, styled):
    gen = torch.Generator().manual_seed(5)
    x = torch.randn(shape, generator=gen) * 3 + 11.0           # large mean: stresses the variance accumulation
    g = torch.randn(shape, generator=gen)
    gain = torch.randn((shape[0] if styled else 1, shape[1], 1, 1), generator=gen)
    bias = torch.randn((1, shape[1], 1, 1), generator=gen)
    xo, go, bo = (t.double().clone().requires_grad_(True) for t in (x, gain, bias))
    yo = O.whole_tensor_norm(xo, go, bo)
    yo.backward(g.double())
    xd, gd, bd = (dev(t).requires_grad_(True) for t in (x, gain, bias))
    yd = ops.whole_norm(xd, gd, bd)
    yd.backward(dev(g))
    close(yd, yo, what="y")
    close(xd.grad, xo.grad, atol_frac=1e-4, what="dx")
    close(gd.grad, go.grad, atol_frac=1e-4, what="dgain")
    close(bd.grad, bo.grad, what="dbias")


def test_gate_golden(golden):
    r = golden("primitives.pt")["gate"]
    x = dev(r["x"]).requires_grad_(True)
    gamma = dev(r["gamma"]).requires_grad_(True)
    y = ops.gate(x, x * 0.5 + 0.1, gamma, True)
    y.backward(dev(r["g"]))
    close(y, r["y"], what="y")
    close(gamma.grad, r["dgamma"], what="dgamma")
    close(x.grad, r["dx"], what="dx")


@pytest.mark.parametrize("tag,build", [
    ("sn_conv", lambda: torch.nn.Conv2d(5, 7, 3, padding=1, bias=False)),
    ("sn_convT", lambda: torch.nn.ConvTranspose2d(5, 7, 4, stride=2, padding=1, bias=False)),
    ("sn_linear", lambda: torch.nn.Linear(9, 4)),
])
def test_spectral_norm_golden(golden, tag, build):
    r = golden("primitives.pt")[tag]
    sn = layers.SpectralNorm(build())
    sn.load_state_dict(r["state"])
    sn = sn.to(DEV)
    x = dev(r["x"]).requires_grad_(True)
    y = sn(x)
    y.backward(dev(r["g"]))
    close(y, r["y"], what="y")
    close(x.grad, r["dx"], what="dx")
    named = dict(sn.named_parameters())
    for k, g in r["grads"].items():
        close(named[k].grad, g, what=k)
    sd = sn.state_dict()
    for k in ("module.weight_u", "module.weight_v"):
        close(sd[k], r["state_after_1"][k], what=k)
    y2 = sn(dev(r["x"]))
    close(y2, r["y_second"], what="y second call")
    sd = sn.state_dict()
    for k in ("module.weight_u", "module.weight_v"):
        close(sd[k], r["state_after_2"][k], what=k + " second")


@pytest.mark.parametrize("ratio", [2, 4])
def test_feature_pool_golden(golden, ratio):
    r = golden("primitives.pt")[f"featpool_r{ratio}"]
    x = dev(r["x"]).requires_grad_(True)
    y = layers.FeaturePooling(8 // ratio)(x)
    y.backward(dev(r["g"]))
    close(y, r["y"], what="y")
    close(x.grad, r["dx"], what="dx")


def _run_module(m, r, prefix=""):
    m.load_state_dict(r["state"])
    m = m.to(DEV)
    x = dev(r["x"]).requires_grad_(True)
    y = m(x)
    y.backward(dev(r["g"]))
    close(y, r["y"], what="y")
    close(x.grad, r["dx"], atol_frac=1e-4, what="dx")
    named = dict(m.named_parameters())
    assert set(r["grads"]) == {k for k, p in named.items() if p.grad is not None}
    for k, g in r["grads"].items():
        close(named[k].grad, g, atol_frac=1e-4, what="grad " + k)
    sd = m.state_dict()
    for k, v in r["state_after"].items():
        close(sd[k], v, what="state " + k)


@pytest.mark.parametrize("tag", ["skip_up_pool", "skip_up_cat", "skip_down_cat", "skip_down_same", "skip_down_conv"])
def test_skip_path_golden(golden, tag):
    r = golden("primitives.pt")[tag]
    m = layers.Scale(*r["args"])
    if not isinstance(m, torch.nn.Module):
        m = torch.nn.Sequential()
    _run_module(m, r)


def test_self_attention_golden(golden):
    _run_module(layers.SelfAttention(8), golden("primitives.pt")["selfattn"])


def test_feature_attention_golden(golden):
    _run_module(layers.feature_attention(4, 8), golden("primitives.pt")["featattn"])


@pytest.mark.parametrize("tag", ["deep_up", "deep_down", "deep_flat", "deep_depth3", "deep_depth4_up"])
def test_deep_conv_golden(golden, tag):
    r = golden("primitives.pt")[tag]
    _run_module(layers.DeepResidualConv(*r["args"]), r)


def test_style_linear_golden(golden):
    r = golden("primitives.pt")["linear"]
    m = layers.LinearModule(7, 5)
    m.load_state_dict(r["state"])
    m = m.to(DEV)
    x = dev(r["x"]).requires_grad_(True)
    act, pre = m(x)
    torch.autograd.backward([act, pre], [dev(r["g_act"]), dev(r["g_pre"])])
    close(act, r["act"], what="act")
    close(pre, r["pre"], what="pre")
    close(x.grad, r["dx"], what="dx")
    named = dict(m.named_parameters())
    for k, g in r["grads"].items():
        close(named[k].grad, g, what=k)


@pytest.mark.parametrize("b,f,hw", [(2, 96, 64), (1, 33, 1000), (3, 64, 16384)])
def test_softmax_pixels_vs_oracle(b, f, hw):
    gen = torch.Generator().manual_seed(7)
    x = torch.randn((b, f, hw, 1), generator=gen) * 3
    g = torch.randn((b, f, hw, 1), generator=gen)
    xo = x.double().requires_grad_(True)
    yo = torch.softmax(xo, dim=2)
    yo.backward(g.double())
    xd = dev(x).requires_grad_(True)
    yd = ops.SoftmaxPixelsFn.apply(xd)
    yd.backward(dev(g))
    close(yd, yo, what="y")
    close(xd.grad, xo.grad, atol_frac=1e-4, what="dx")
    assert torch.allclose(yd.sum(dim=2).cpu(), torch.ones(b, f, 1), atol=1e-5)


@pytest.mark.parametrize("cfg", [
    dict(kind="conv", cin=64, cout=96, k=5, s=2, p=2, h=16, b=3),
    dict(kind="conv", cin=3, cout=3, k=5, s=2, p=2, h=32, b=2),
    dict(kind="conv", cin=48, cout=3, k=1, s=1, p=0, h=16, b=2),
    dict(kind="conv", cin=40, cout=24, k=3, s=1, p=1, h=9, b=2),
    dict(kind="convT", cin=96, cout=96, k=4, s=2, p=1, h=8, b=2),
    dict(kind="convT", cin=20, cout=52, k=1, s=1, p=0, h=5, b=3),
    dict(kind="conv", cin=32, cout=16, k=(8, 1), s=1, p=0, h=8, b=4),
])
def test_conv_family_vs_oracle(cfg):
    import torch.nn.functional as F
    gen = torch.Generator().manual_seed(11)
    k = cfg["k"] if isinstance(cfg["k"], tuple) else (cfg["k"], cfg["k"])
    if cfg["kind"] == "conv":
        mod = torch.nn.Conv2d(cfg["cin"], cfg["cout"], k, cfg["s"], cfg["p"], bias=False)
    else:
        mod = torch.nn.ConvTranspose2d(cfg["cin"], cfg["cout"], k, cfg["s"], cfg["p"], bias=False)
    sn = layers.SpectralNorm(mod)
    state = {kk: v.clone() for kk, v in sn.state_dict().items()}
    x = torch.randn((cfg["b"], cfg["cin"], cfg["h"], cfg["h"]), generator=gen)
    st = O.load_state({kk.replace("module.", "", 1): v.double() for kk, v in state.items()})
    xo = x.double().requires_grad_(True)
    w = O.power_iterate(st, "")
    yo = F.conv2d(xo, w, None, cfg["s"], cfg["p"]) if cfg["kind"] == "conv" else F.conv_transpose2d(xo, w, None, cfg["s"], cfg["p"])
    g = torch.randn(yo.shape, generator=gen)
    yo.backward(g.double())
    sn = sn.to(DEV)
    xd = dev(x).requires_grad_(True)
    yd = sn(xd)
    yd.backward(dev(g))
    close(yd, yo, atol_frac=1e-4, what="y")
    close(xd.grad, xo.grad, atol_frac=1e-4, what="dx")
    close(sn.module.weight_bar.grad, st["weight_bar"].grad, atol_frac=2e-4, what="dW")
    close(sn.module.weight_u, st["weight_u"], what="u")
    close(sn.module.weight_v, st["weight_v"], what="v")


# ------------------------------------------------------------------------------------------ full models
def _build(rec):
    L.configure(**rec["overrides"])
    torch.manual_seed(999)
    gen, g_opt = L.get_model(L.Generator(), L.CFG.GLR, DEV)
    dis, d_opt = L.get_model(L.Discriminator(), L.CFG.DLR, DEV)
    gen.load_state_dict(rec["g_state"])
    dis.load_state_dict(rec["d_state"])
    gen.noise = dev(rec["const_noise"])
    return gen, dis, g_opt, d_opt


@pytest.mark.parametrize("name", ["step_s32_w2_b3.pt", "step_s16_w2_depth3_b2.pt", "step_s16_w4_separable_b2.pt"])
def test_full_forward_golden(golden, name):
    r = golden(name)
    gen, dis, _, _ = _build(r)
    with torch.no_grad():
        img = gen(dev(r["z"]))
        logit = dis(dev(r["real"]))
    close(img, r["g_out"], atol_frac=1e-4, what="G(z)")
    close(logit, r["d_out_real"], atol_frac=1e-4, what="D(real)")
    sd = gen.state_dict()
    for k, v in r["g_state_after_fwd"].items():
        close(sd[k], v, what=k)


@pytest.mark.parametrize("name", ["step_s32_w2_b3.pt", "step_s16_w2_depth3_b2.pt", "step_s16_w4_separable_b2.pt"])
def test_full_training_step_golden(golden, name):
    r = golden(name)
    gen, dis, g_opt, d_opt = _build(r)
    trainer = L.GanTrainer(gen, dis, g_opt, d_opt)
    real, aug, z = dev(r["real"]), dev(r["aug"]), dev(r["z"])

    grads = {}
    orig = {}
    for tag, opt, model in (("d", d_opt, dis), ("g", g_opt, gen)):
        orig[tag] = opt.step

        def spy(closure=None, tag=tag, opt=opt, model=model):
            grads[tag] = {k: p.grad.detach().clone() for k, p in model.named_parameters() if p.requires_grad}
            return orig[tag]()
        opt.step = spy
    d_out, g_out = trainer.step(real, aug, z)
    close(d_out[0], r["d_error"], what="d hinge loss")
    close(d_out[1], r["penalty"], rtol=2e-3, what="penalty")
    close(g_out[0], r["g_error"], rtol=1e-3, what="g loss")
    for k, g in r["d_grads"].items():
        close(grads["d"][k], g, rtol=2e-3, atol_frac=5e-4, what="D grad " + k)
    for k, g in r["g_grads"].items():
        close(grads["g"][k], g, rtol=2e-3, atol_frac=5e-4, what="G grad " + k)
    # parameters the reference never gives a gradient (styled norm weights) must stay untouched
    for k, p in gen.named_parameters():
        if p.requires_grad and k not in r["g_grads"]:
            assert float(grads["g"][k].abs().max()) == 0.0, k
    dsd, gsd = dis.state_dict(), gen.state_dict()
    for k, v in r["d_state_after_opt"].items():
        if not k.endswith(("_u", "_v")):
            frac = ((dsd[k].cpu() - v).abs() > 1e-5 + 1e-3 * v.abs()).float().mean().item()
            assert frac < 2e-3, f"D param {k}: {frac:.2e} of elements moved differently"
    for k, v in r["g_state_after_opt"].items():
        tol = 1e-3 if k.endswith(("_u", "_v")) else 1e-5
        frac = ((gsd[k].cpu() - v).abs() > tol + 1e-3 * v.abs()).float().mean().item()
        assert frac < 2e-3, f"G param {k}: {frac:.2e} of elements moved differently"
    for k, v in r["d_state_final"].items():
        close(dsd[k], v, rtol=2e-3, atol_frac=2e-3, what="D final " + k)


def test_start_layer_forward_fails_like_the_reference(golden):
    """START_LAYER = 1: the reference's forward raises a shape error in the style chain's first Linear (fixture key
    `forward_error`, generated by running the reference); the drop-in raises at the same layer instead of inventing
    semantics the reference does not have."""
    r = golden("start_layer1_s16_w2.pt")
    assert "cannot be multiplied" in r["forward_error"]
    L.configure(**r["overrides"])
    gen, _ = L.get_model(L.Generator(), L.CFG.GLR, DEV)
    with pytest.raises((ValueError, RuntimeError), match="input channels|shape"):
        gen(torch.randn(2, L.CFG.INPUT_VECTOR_Z, device=DEV))


def test_three_training_steps_golden(golden):
    """Three consecutive steps against the REAL reference (tests/golden/steps3_*.pt): from the second discriminator step
    on the reference trains D's weight_u / weight_v (main.py:172 switches them on, sigma = u.(W v) at
    spectral_norm.py:31 hands them gradients, Nadam holds them since utils.py:149) -- losses, the u / v gradients and
    the u / v themselves after every step must follow."""
    r = golden("steps3_s16_w2_b2.pt")
    gen, dis, g_opt, d_opt = _build(r)
    trainer = L.GanTrainer(gen, dis, g_opt, d_opt)
    real, aug, z = dev(r["real"]), dev(r["aug"]), dev(r["z"])
    uv = {}
    orig = d_opt.step

    def spy(closure=None):
        dis._finish_uv_grads()
        uv.clear()
        uv.update({k: p.grad.detach().clone() for k, p in dis.named_parameters()
                   if k.endswith(("_u", "_v")) and p.requires_grad and p.grad is not None})
        return orig()
    d_opt.step = spy
    for step in range(r["steps"]):
        d_out, g_out = trainer.step(real, aug, z)
        want = r["losses"][step]
        assert abs(d_out[0].item() - want[0]) < 1e-3 * abs(want[0]), (step, d_out[0].item(), want)
        assert abs(d_out[1].item() - want[1]) < 2e-2 * abs(want[1]) + 1e-7, (step, d_out[1].item(), want)
        assert abs(g_out[0].item() - want[2]) < 2e-3 * abs(want[2]), (step, g_out[0].item(), want)
        assert set(uv) == set(r["d_uv_grads"][step]), (step, sorted(uv))
        for k, g in r["d_uv_grads"][step].items():
            close(uv[k], g, rtol=1e-2, atol_frac=1e-2, what=f"step {step} grad {k}")
        dsd = dis.state_dict()
        for k, v in r["d_state_per_step"][step].items():
            if k.endswith(("_u", "_v")):
                close(dsd[k], v, rtol=1e-2, atol_frac=1e-2, what=f"step {step} {k}")
    assert len(r["d_uv_grads"][1]) > 0


def test_forward_is_bit_reproducible(golden):
    """The reference is deterministic (plain torch.mv / conv on CPU): the same state must give the same images, logits
    and u / v, run after run and model instance after model instance (parameters re-homed in the optimizer arena or
    freshly allocated).  No floating-point atomics on the forward path: spectral-norm partial sums, split-K partial
    tiles and the norm statistics are all reduced in a fixed order."""
    r = golden("step_s32_w2_b3.pt")
    z, real = dev(r["z"]), dev(r["real"])
    outs = []
    for trial in range(3):
        L.config.reset()
        L.configure(PRECISION="bf16", **r["overrides"])   # the product (tensor-core) path, split-K layers included
        gen, dis = L.Generator().to(DEV), L.Discriminator().to(DEV)
        if trial == 1:                                   # parameters as views of the flat Nadam arena
            L.Nadam(gen.parameters(), lr=1e-3)
            L.Nadam(dis.parameters(), lr=1e-3)
        gen.load_state_dict(r["g_state"])
        dis.load_state_dict(r["d_state"])
        gen.noise = dev(r["const_noise"])
        with torch.no_grad():
            img = gen(z)
            logit = dis(real)
            img2 = gen(z)                                # second call: u / v have advanced, as in the reference
        torch.cuda.synchronize()
        outs.append((img.clone(), logit.clone(), img2.clone(),
                     {k: v.clone() for k, v in gen.state_dict().items() if k.endswith(("_u", "_v"))},
                     {k: v.clone() for k, v in dis.state_dict().items() if k.endswith(("_u", "_v"))}))
    for other in outs[1:]:
        assert torch.equal(outs[0][0], other[0]), "G(z) differs between runs"
        assert torch.equal(outs[0][1], other[1]), "D(x) differs between runs"
        assert torch.equal(outs[0][2], other[2]), "second G(z) differs between runs"
        for a, b in ((outs[0][3], other[3]), (outs[0][4], other[4])):
            for k in a:
                assert torch.equal(a[k], b[k]), k
    assert not torch.equal(outs[0][0], outs[0][2])       # every forward runs a power iteration (spectral_norm.py:57-59)


def test_step_vs_oracle_default_width():
    """32x32 with the reference's default channel widths, batch 4, vs the CPU oracle."""
    L.configure(IMAGE_SIZE=32)
    cfg = O.OracleConfig(IMAGE_SIZE=32)
    torch.manual_seed(999)
    gen, g_opt = L.get_model(L.Generator(), L.CFG.GLR, DEV)
    dis, d_opt = L.get_model(L.Discriminator(), L.CFG.DLR, DEV)
    gs = O.load_state({k: v.cpu() for k, v in gen.state_dict().items()})
    ds = O.load_state({k: v.cpu() for k, v in dis.state_dict().items()})
    real, aug, z = O.synthetic_batch(cfg, 4)
    o_d, o_pen, o_g = O.train_step(gs, ds, gen.noise.cpu(), real, aug, z, cfg,
                                   O.Nadam(cfg.GLR, (cfg.BETA_1, cfg.BETA_2)), O.Nadam(cfg.DLR, (cfg.BETA_1, cfg.BETA_2)))
    d_out, g_out = L.GanTrainer(gen, dis, g_opt, d_opt).step(dev(real), dev(aug), dev(z))
    close(d_out[0], o_d, rtol=1e-3, what="d loss")
    close(d_out[1], o_pen, rtol=5e-3, what="penalty")
    close(g_out[0], o_g, rtol=2e-3, what="g loss")
    for k, p in gen.named_parameters():
        if p.requires_grad and gs[k].grad is not None:
            close(p.grad, gs[k].grad, rtol=5e-3, atol_frac=1e-3, what="G grad " + k)


@pytest.mark.parametrize("b,c,h,w", [(3, 3, 16, 12), (2, 1, 8, 8), (2, 40, 9, 7), (5, 8, 4, 4)])
def test_layout_round_trip(b, c, h, w):
    """NCHW <-> channels-last at the model boundary (few-channel and tiled-transpose kernels)."""
    x = torch.randn((b, c, h, w), generator=torch.Generator().manual_seed(3)).to(DEV)
    cl = ops._to_channels_last(x)
    assert cl.shape == x.shape and torch.equal(cl, x)                      # same logical tensor
    assert torch.equal(cl.permute(0, 2, 3, 1).contiguous(), x.permute(0, 2, 3, 1).contiguous())
    back = ops.to_nchw(cl)
    assert back.is_contiguous() and torch.equal(back, x)


def test_programmatic_dependent_launch_is_bit_identical_to_serialised_launches(golden):
    """Every kernel of the library is launched with the programmatic-stream-serialization attribute and waits for its
    predecessor on the device (griddepcontrol.wait) before touching global memory.  A kernel that read or wrote anything
    ahead of that wait would race with its predecessor; the forward pass is deterministic, so the two launch modes must
    agree bit for bit (images, logits, power-iteration vectors); the backward pass (whose norm / gate column sums use
    float atomics) must agree to rounding noise."""
    from locate_b200 import _lib
    r = golden("step_s32_w2_b3.pt")
    z, real = dev(r["z"]), dev(r["real"])
    outs = []
    was = _lib.set_pdl(True)
    try:
        for pdl in (False, True, True):
            _lib.set_pdl(pdl)
            L.config.reset()
            L.configure(PRECISION="bf16", **r["overrides"])
            gen, dis = L.Generator().to(DEV), L.Discriminator().to(DEV)
            gen.load_state_dict(r["g_state"])
            dis.load_state_dict(r["d_state"])
            gen.noise = dev(r["const_noise"])
            img = gen(z)
            logit = dis(img)
            logit.mean().backward()
            torch.cuda.synchronize()
            big = [p.grad.clone() for p in list(gen.parameters()) + list(dis.parameters())
                   if p.grad is not None and p.dim() >= 2 and min(p.shape[:2]) > 4]
            outs.append((img.detach().clone(), logit.detach().clone(), big,
                         {k: v.clone() for k, v in dis.state_dict().items() if k.endswith(("_u", "_v"))}))
    finally:
        _lib.set_pdl(was)
    for other in outs[1:]:
        assert torch.equal(outs[0][0], other[0]) and torch.equal(outs[0][1], other[1])
        assert all(torch.equal(a, b) for a, b in zip(outs[0][3].values(), other[3].values()))
        assert len(other[2]) == len(outs[0][2]) > 10
        for a, b in zip(outs[0][2], other[2]):
            assert float((a - b).norm()) <= 1e-3 * float(a.norm()) + 1e-12
