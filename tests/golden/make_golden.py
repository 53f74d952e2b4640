"""Generate the committed golden fixtures by RUNNING THE UNMODIFIED REFERENCE (build container only).

    python tests/golden/make_golden.py        # needs /root/reference; writes tests/golden/*.pt

Every fixture holds the seeded inputs, the reference's state_dict BEFORE the call, and what the
reference produced (outputs, gradients, post-forward u/v, post-Nadam parameters).  The fixtures pin
oracle/locate_oracle.py (tests/test_oracle_golden.py) and travel to the GPU box, where
/root/reference does not exist.  torch version recorded in each file.
"""
import copy
import os
import sys
import warnings

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import ref_loader  # noqa: E402

warnings.filterwarnings("ignore")
torch.set_num_threads(4)


def sd(module):
    return {k: v.detach().clone() for k, v in module.state_dict().items()}


def grads(module):
    return {k: p.grad.detach().clone() for k, p in module.named_parameters() if p.grad is not None}


def save(name, payload):
    payload["torch_version"] = str(torch.__version__)
    path = os.path.join(HERE, name)
    torch.save(payload, path)
    print(f"wrote {name}: {os.path.getsize(path) / 1e6:.2f} MB")


def randn(*shape, seed):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed))


def primitives(libs):
    from libs.activation import nonlinear_function
    from libs.inplace_norm import InPlaceNorm
    from libs.merge import ResModule
    from libs.scale import FeaturePooling, Scale
    from libs.attention import SelfAttention, feature_attention
    from libs.conv import DeepResidualConv
    from libs.spectral_norm import SpectralNorm
    from libs.linear import LinearModule
    out = {}

    # RootTanh incl. the |x| > 44 overflow region (activation.py:24-26)
    x = torch.cat([randn(257, seed=10) * 3, torch.tensor([0.0, -0.0, 44.0, -50.0, 90.0, 400.0, -400.0, 1e-4])])
    x.requires_grad_(True)
    y = nonlinear_function(x)
    g = randn(*x.shape, seed=11)
    y.backward(g)
    out["roottanh"] = dict(x=x.detach().clone(), y=y.detach().clone(), g=g, dx=x.grad.clone())

    # whole-tensor norm, per-channel gain and per-sample style gain (inplace_norm.py:40-45)
    for tag, styled in (("norm_plain", False), ("norm_styled", True)):
        torch.manual_seed(20)
        m = InPlaceNorm(6)
        m.weight.data.uniform_(0.5, 1.5)
        m.bias.data.uniform_(-0.5, 0.5)
        x = (randn(3, 6, 5, 7, seed=21) * 2 + 0.7).requires_grad_(True)
        scale = (randn(3, 6, 1, 1, seed=22)).requires_grad_(True) if styled else None
        y = m(x, scale)
        g = randn(*y.shape, seed=23)
        y.backward(g)
        rec = dict(x=x.detach().clone(), weight=m.weight.data.clone(), bias=m.bias.data.clone(), y=y.detach().clone(),
                   g=g, dx=x.grad.clone(), dbias=m.bias.grad.clone())
        if styled:
            rec.update(scale=scale.detach().clone(), dscale=scale.grad.clone())
            assert m.weight.grad is None
        else:
            rec.update(dweight=m.weight.grad.clone())
        out[tag] = rec

    # gated residual incl. the d-gamma quirk (merge.py:31-39)
    torch.manual_seed(30)
    m = ResModule(lambda t: t, lambda t: t * 0.5 + 0.1, m=3)
    x = randn(2, 4, 3, 3, seed=31).requires_grad_(True)
    y = m(x)
    g = randn(*y.shape, seed=32)
    y.backward(g)
    out["gate"] = dict(x=x.detach().clone(), gamma=m.gamma.data.clone(), y=y.detach().clone(), g=g,
                       dx=x.grad.clone(), dgamma=m.gamma.grad.clone())

    # spectral norm: conv weight, conv-transpose weight, linear (spectral_norm.py:21-32)
    for tag, mod, inp in (("sn_conv", torch.nn.Conv2d(5, 7, 3, padding=1, bias=False), randn(2, 5, 6, 6, seed=41)),
                          ("sn_convT", torch.nn.ConvTranspose2d(5, 7, 4, stride=2, padding=1, bias=False), randn(2, 5, 3, 3, seed=42)),
                          ("sn_linear", torch.nn.Linear(9, 4), randn(3, 9, seed=43))):
        torch.manual_seed(40)
        sn = SpectralNorm(mod)
        before = sd(sn)
        inp = inp.requires_grad_(True)
        y1 = sn(inp)
        g = randn(*y1.shape, seed=44)
        y1.backward(g)
        mid = sd(sn)
        y2 = sn(inp.detach())           # second call: u/v advance again
        out[tag] = dict(state=before, x=inp.detach().clone(), y=y1.detach().clone(), g=g, dx=inp.grad.clone(),
                        grads=grads(sn), state_after_1=mid, y_second=y2.detach().clone(), state_after_2=sd(sn))

    # FeaturePooling (scale.py:12-16), ratios 2 and 4
    for r in (2, 4):
        x = randn(2, 8, 4, 6, seed=50 + r).requires_grad_(True)
        y = FeaturePooling(8 // r)(x)
        g = randn(*y.shape, seed=52)
        y.backward(g)
        out[f"featpool_r{r}"] = dict(x=x.detach().clone(), y=y.detach().clone(), g=g, dx=x.grad.clone())

    # skip paths: up (pool+bilinear, cat+bilinear), down (cat+avgpool, avgpool only, non-divisible conv)
    for tag, args, shape in (("skip_up_pool", (8, 4, 2, True), (2, 8, 4, 4)),
                             ("skip_up_cat", (4, 10, 2, True), (2, 4, 2, 2)),
                             ("skip_down_cat", (4, 8, 2, False), (2, 4, 8, 8)),
                             ("skip_down_same", (6, 6, 2, False), (2, 6, 4, 4)),
                             ("skip_down_conv", (6, 4, 2, False), (2, 6, 4, 4))):
        torch.manual_seed(60)
        m = Scale(*args)
        is_mod = isinstance(m, torch.nn.Module)
        before = sd(m) if is_mod else {}
        x = randn(*shape, seed=61).requires_grad_(True)
        y = m(x)
        g = randn(*y.shape, seed=62)
        y.backward(g)
        out[tag] = dict(args=args, state=before, x=x.detach().clone(), y=y.detach().clone(), g=g, dx=x.grad.clone(),
                        grads=grads(m) if is_mod else {}, state_after=sd(m) if is_mod else {})

    # self attention / feature attention (attention.py:9-54)
    torch.manual_seed(70)
    m = SelfAttention(8)
    before = sd(m)
    x = randn(2, 8, 4, 4, seed=71).requires_grad_(True)
    y = m(x)
    g = randn(*y.shape, seed=72)
    y.backward(g)
    out["selfattn"] = dict(state=before, x=x.detach().clone(), y=y.detach().clone(), g=g, dx=x.grad.clone(),
                           grads=grads(m), state_after=sd(m))
    torch.manual_seed(73)
    m = feature_attention(4, 8)
    before = sd(m)
    x = randn(3, 8, 4, 4, seed=74).requires_grad_(True)
    y = m(x)
    g = randn(*y.shape, seed=75)
    y.backward(g)
    out["featattn"] = dict(state=before, x=x.detach().clone(), y=y.detach().clone(), g=g, dx=x.grad.clone(),
                           grads=grads(m), state_after=sd(m))

    # DeepResidualConv variants (conv.py:27-72): transposed 4x4 s2, 5x5 s2, 3x3 s1, depth 3 bottleneck
    for tag, args, shape in (("deep_up", (8, 4, True, 2), (2, 8, 3, 3)),
                             ("deep_down", (4, 8, False, 2), (2, 4, 6, 6)),
                             ("deep_flat", (6, 3, False, 1, False, 2, 1), (2, 6, 5, 5)),
                             ("deep_depth3", (8, 8, False, 2, True, 2, 3), (2, 8, 8, 8)),
                             ("deep_depth4_up", (16, 8, True, 2, True, 2, 4), (2, 16, 3, 3))):
        torch.manual_seed(80)
        m = DeepResidualConv(*args)
        before = sd(m)
        x = randn(*shape, seed=81).requires_grad_(True)
        y = m(x)
        g = randn(*y.shape, seed=82)
        y.backward(g)
        out[tag] = dict(args=args, state=before, x=x.detach().clone(), y=y.detach().clone(), g=g, dx=x.grad.clone(),
                        grads=grads(m), state_after=sd(m))

    # style linear (linear.py:13-15)
    torch.manual_seed(90)
    m = LinearModule(7, 5)
    before = sd(m)
    x = randn(3, 7, seed=91).requires_grad_(True)
    act, pre = m(x)
    g1, g2 = randn(*act.shape, seed=92), randn(*pre.shape, seed=93)
    (act * g1).sum().add((pre * g2).sum()).backward()
    out["linear"] = dict(state=before, x=x.detach().clone(), act=act.detach().clone(), pre=pre.detach().clone(),
                         g_act=g1, g_pre=g2, dx=x.grad.clone(), grads=grads(m), state_after=sd(m))
    return out


def full_model(tag, batch, **overrides):
    """One G forward, one D forward and ONE full training step of the reference
    (main.py:142-172 with miniter = MINIBATCHES = DITERS = 1)."""
    libs = ref_loader.load_ref(**overrides)
    cfg = sys.modules["libs.config"]
    with ref_loader.quiet():
        torch.manual_seed(999)
        gen, g_opt = libs.get_model(libs.Generator(), cfg.GLR, cfg.DEVICE)
        dis, d_opt = libs.get_model(libs.Discriminator(), cfg.DLR, cfg.DEVICE)
    s, zdim = cfg.IMAGE_SIZE, cfg.INPUT_VECTOR_Z
    real = randn(batch, 3, s, s, seed=0).clamp(-1, 1)
    aug = (real + 0.05 * randn(batch, 3, s, s, seed=1)).clamp(-1, 1)
    z = randn(batch, zdim, seed=2)
    rec = dict(overrides=overrides, batch=batch, real=real, aug=aug, z=z, const_noise=gen.noise.clone(),
               g_state=sd(gen), d_state=sd(dis))

    # plain forwards on deep copies (each forward mutates u/v)
    g2, d2 = copy.deepcopy(gen), copy.deepcopy(dis)
    g2.noise = gen.noise
    with torch.no_grad():
        rec["g_out"] = g2(z).clone()
        rec["d_out_real"] = d2(real).clone()
    rec["g_state_after_fwd"] = {k: v for k, v in sd(g2).items() if k.endswith(("_u", "_v"))}

    # the step
    generated = gen(z).detach()
    dis.zero_grad()
    d_true = dis(real).view(-1)
    d_gen = -dis(generated).view(-1)
    d_error = (libs.hinge(d_true) + libs.hinge(d_gen)).mean()
    pen = libs.penalty(d_true, aug, dis, cfg.DEVICE)
    (d_error + pen).backward()
    rec.update(fake=generated.clone(), d_true=d_true.detach().clone(), d_gen=d_gen.detach().clone(),
               d_error=d_error.detach().clone(), penalty=pen.detach().clone(), d_grads=grads(dis))
    d_opt.step()
    rec["d_state_after_opt"] = sd(dis)
    dis.requires_grad_(False)
    gen.zero_grad()
    g_error = libs.hinge(dis(gen(z)).view(-1)).mean()
    g_error.backward()
    rec.update(g_error=g_error.detach().clone(), g_grads=grads(gen))
    g_opt.step()
    dis.requires_grad_(True)
    rec["g_state_after_opt"] = sd(gen)
    rec["d_state_final"] = {k: v for k, v in sd(dis).items() if k.endswith(("_u", "_v"))}
    save(f"{tag}.pt", rec)


def multi_step(tag, batch, steps, **overrides):
    """`steps` consecutive training steps of the reference (main.py:142-172, miniter = MINIBATCHES = DITERS = 1) on the
    same batch.  Pins what a single step cannot: `dis.requires_grad_(True)` at main.py:172 also switches on the
    discriminator's weight_u / weight_v (requires_grad=False Parameters, spectral_norm.py:45-46), so from the SECOND
    discriminator step on they receive gradients through sigma = u.(W v) and Nadam -- whose parameter list holds them
    since construction (utils.py:149) -- moves them, each with its own step counter / momentum schedule."""
    libs = ref_loader.load_ref(**overrides)
    cfg = sys.modules["libs.config"]
    with ref_loader.quiet():
        torch.manual_seed(999)
        gen, g_opt = libs.get_model(libs.Generator(), cfg.GLR, cfg.DEVICE)
        dis, d_opt = libs.get_model(libs.Discriminator(), cfg.DLR, cfg.DEVICE)
    s, zdim = cfg.IMAGE_SIZE, cfg.INPUT_VECTOR_Z
    real = randn(batch, 3, s, s, seed=0).clamp(-1, 1)
    aug = (real + 0.05 * randn(batch, 3, s, s, seed=1)).clamp(-1, 1)
    z = randn(batch, zdim, seed=2)
    rec = dict(overrides=overrides, batch=batch, steps=steps, real=real, aug=aug, z=z, const_noise=gen.noise.clone(),
               g_state=sd(gen), d_state=sd(dis), losses=[], d_uv_grads=[], d_state_per_step=[], g_state_per_step=[])
    for _ in range(steps):
        generated = gen(z).detach()
        dis.zero_grad()
        d_true = dis(real).view(-1)
        d_gen = -dis(generated).view(-1)
        d_error = (libs.hinge(d_true) + libs.hinge(d_gen)).mean()
        pen = libs.penalty(d_true, aug, dis, cfg.DEVICE)
        (d_error + pen).backward()
        rec["d_uv_grads"].append({k: g for k, g in grads(dis).items() if k.endswith(("_u", "_v"))})
        d_opt.step()
        dis.requires_grad_(False)
        gen.zero_grad()
        g_error = libs.hinge(dis(gen(z)).view(-1)).mean()
        g_error.backward()
        g_opt.step()
        dis.requires_grad_(True)
        rec["losses"].append((d_error.item(), pen.item(), g_error.item()))
        rec["d_state_per_step"].append(sd(dis))
        rec["g_state_per_step"].append(sd(gen))
    save(f"{tag}.pt", rec)


def start_layer(tag, **overrides):
    """START_LAYER >= 1 (models.py:44-50): what the reference CONSTRUCTS (state_dict names and shapes) and what its own
    forward does with it -- it raises: the style chain's first Linear is built for the input block's output width but is
    fed the Z-dimensional latent (block.py:88-96)."""
    libs = ref_loader.load_ref(**overrides)
    with ref_loader.quiet():
        torch.manual_seed(999)
        gen = libs.Generator()
    rec = dict(overrides=overrides, g_shapes={k: tuple(v.shape) for k, v in gen.state_dict().items()})
    try:
        gen(randn(2, sys.modules["libs.config"].INPUT_VECTOR_Z, seed=2))
        rec["forward_error"] = None
    except RuntimeError as exc:
        rec["forward_error"] = str(exc)
    print("START_LAYER forward:", rec["forward_error"])
    save(f"{tag}.pt", rec)


if __name__ == "__main__":
    libs = ref_loader.load_ref()
    save("primitives.pt", primitives(libs))
    full_model("step_s32_w2_b3", 3, IMAGE_SIZE=32, BASE_FEATURE_FACTOR=2)
    full_model("step_s16_w2_depth3_b2", 2, IMAGE_SIZE=16, BASE_FEATURE_FACTOR=2, DEPTH=3)
    multi_step("steps3_s16_w2_b2", 2, 3, IMAGE_SIZE=16, BASE_FEATURE_FACTOR=2)
    # SEPARABLE = True (config.py:53): depthwise k x k convs (conv.py:17) and the grouped full-extent feature-attention
    # conv (attention.py:15-21); BASE_FEATURE_FACTOR = 4 so that both attention sizes (8, 16 / 8) keep >= 1 group
    full_model("step_s16_w4_separable_b2", 2, IMAGE_SIZE=16, BASE_FEATURE_FACTOR=4, SEPARABLE=True)
    start_layer("start_layer1_s16_w2", IMAGE_SIZE=16, BASE_FEATURE_FACTOR=2, START_LAYER=1)
