"""CPU ORACLE for the LocAtE GAN-training hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may import this file; the product package `locate_b200` never does.

What it is: an independent pure-torch (CPU, fp32 or fp64) *functional* restatement of the
arithmetic the reference performs on the generator/discriminator forward+backward path.  There is
no nn.Module tree here: every function takes a flat ``state`` dict whose keys are the
reference's ``state_dict`` names, so a reference checkpoint (or a golden fixture exported from
the reference) plugs in directly.

Parity pinning: the reference ships NO tests / golden vectors, so the oracle is
pinned against OUTPUTS OF THE REFERENCE ITSELF run in the build container:
``tests/golden/make_golden.py`` imports /root/reference, runs it on seeded inputs and commits
inputs + state + outputs + gradients under ``tests/golden/*.pt``; ``tests/test_oracle_golden.py``
replays them through this file.  The arithmetic below the Python (conv/matmul/softmax) is
torch 2.11.0+cu128 CPU (ATen/oneDNN) -- the same library the reference calls, unpinned by the
reference (no requirements file).

Each function cites the reference lines it restates (paths relative to /root/reference).
"""
from __future__ import annotations

import dataclasses
import math
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

State = Dict[str, torch.Tensor]


# --------------------------------------------------------------------------------------
# configuration (libs/config.py:19-73) -- same field names, but a value object
# --------------------------------------------------------------------------------------
@dataclasses.dataclass(frozen=True)
class OracleConfig:
    IMAGE_SIZE: int = 128
    FACTOR: int = 2
    G_STRIDE: int = 2
    D_STRIDE: int = 2
    BASE_FEATURE_FACTOR: int = 8
    BOTTLENECK: int = 4
    MIN_ATTENTION_SIZE: int = 8
    ATTENTION_EVERY_NTH_LAYER: int = 2
    DEPTH: int = 1
    SEPARABLE: bool = False            # config.py:53: depthwise k x k convs (conv.py:17), grouped full-extent feature attention
    START_LAYER: int = 0               # config.py:39: depth of the generator's input block (models.py:46-48)
    ROOTTANH_GROWTH: int = 4
    GLR: float = 5e-4
    DLR: float = 2e-3
    BETA_1: float = 0.5
    BETA_2: float = 0.9

    @property
    def LAYERS(self) -> int:  # config.py:50
        return int(math.log(self.IMAGE_SIZE, 2))

    @property
    def INPUT_VECTOR_Z(self) -> int:  # config.py:65
        return self.IMAGE_SIZE

    @property
    def GEN_FEATURES(self) -> int:  # config.py:60
        return self.FACTOR ** int(math.log(self.IMAGE_SIZE, self.G_STRIDE)) * self.BASE_FEATURE_FACTOR * 3

    @property
    def DIS_FEATURES(self) -> int:  # config.py:61
        return self.FACTOR ** int(math.log(self.IMAGE_SIZE, self.D_STRIDE)) * self.BASE_FEATURE_FACTOR


def _quad(n: int) -> int:  # models.py:12-13
    return n // 4 * 4


def generator_features(cfg: OracleConfig) -> List[int]:
    """[in, f_0, ..., f_{L-2}]: models.py:16-22,37-52; in = Z, or with an input block (START_LAYER >= 1) its output
    width GFeatures(0, L-1)(L-1) = quadnorm(GEN_FEATURES)."""
    n = cfg.LAYERS - 1
    widths = [_quad(int(cfg.GEN_FEATURES * cfg.FACTOR ** (i - n))) for i in range(n - 1, -1, -1)]
    first = _quad(int(cfg.GEN_FEATURES)) if cfg.START_LAYER >= 1 else cfg.INPUT_VECTOR_Z
    return [first] + widths


def discriminator_features(cfg: OracleConfig) -> List[int]:
    """[d_0, ..., d_{L-2}, d_{L-2}]: models.py:25-31,72-78."""
    n = cfg.LAYERS - 1
    widths = [_quad(int(cfg.DIS_FEATURES * cfg.FACTOR ** ((i + 1) - n))) for i in range(n)]
    return widths + [widths[-1]]


def _has_attention(cfg: OracleConfig, out_size: int, block_number: int) -> bool:
    """block.py:28-29 (`in_size` there is the block's OUTPUT size, block.py:66-72)."""
    return out_size >= cfg.MIN_ATTENTION_SIZE and block_number % cfg.ATTENTION_EVERY_NTH_LAYER == 0


# --------------------------------------------------------------------------------------
# primitives with the reference's hand-written backward passes
# --------------------------------------------------------------------------------------
class _RootTanh(torch.autograd.Function):
    """y = (x^2+1)^(1/g) * tanh(x)   (activation.py:9-16).

    Backward restates activation.py:20-36, which is exact only for g = 4 (the literal 2 is g/2):
        dy/dx = (2*(x^2+1)/cosh(x)^2 + x*tanh(x)) / (2*(x^2+1)^((g-1)/g))
    cosh^2 overflows to inf for |x| >~ 44 in fp32 and its reciprocal becomes 0 -- kept as is.
    """

    @staticmethod
    def forward(ctx, x, growth):
        ctx.save_for_backward(x)
        ctx.growth = growth
        return (x * x + 1).pow(1.0 / growth) * torch.tanh(x)

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        growth = ctx.growth
        q = x * x + 1
        sech2 = 1.0 / torch.cosh(x).pow(2)
        num = 2 * q * sech2 + x * torch.tanh(x)
        den = 2 * q.pow((growth - 1) / growth)
        return g * num / den, None


def roottanh(x: torch.Tensor, growth: int = 4) -> torch.Tensor:
    return _RootTanh.apply(x, growth)


class _WholeTensorNorm(torch.autograd.Function):
    """out = (x - mean(x)) * gain / std(x) + bias, ONE mean and ONE unbiased std over the whole
    tensor (inplace_norm.py:7-13,40-45).  gain is [1,C,1,1] or a per-sample style [B,C,1,1].

    Backward = inplace_norm.py:17-27 composed with autograd of `x.std()`:
        dx   = gain*g/z - mean(gain*g/z) + dz * (x-mu)/((N-1)*z),   dz = -sum((x-mu)*g*gain)/z^2
        dgain= sum_to_shape((x-mu)*g/z),   dbias = sum_to_shape(g)
    """

    @staticmethod
    def forward(ctx, x, gain, bias):
        mu = x.mean()
        z = x.std()
        ctx.save_for_backward(x, gain, mu, z)
        ctx.bias_shape = bias.shape
        return (x - mu) * gain / z + bias

    @staticmethod
    def backward(ctx, g):
        x, gain, mu, z = ctx.saved_tensors
        n = x.numel()
        xc = x - mu
        direct = gain * g / z
        dz = -(xc * g * gain).sum() / (z * z)
        dx = direct - direct.mean() + dz * xc / ((n - 1) * z)
        dgain = (xc * g / z).sum_to_size(gain.shape)
        dbias = g.sum_to_size(ctx.bias_shape)
        return dx, dgain, dbias


def whole_tensor_norm(x, gain, bias):
    return _WholeTensorNorm.apply(x, gain, bias)


class _Gate(torch.autograd.Function):
    """out = (gamma*y + 1) * x with scalar gamma (merge.py:19-28,61-62).

    Backward restates merge.py:31-39 INCLUDING its quirk: the gamma gradient is sum(x*x*g)
    (the reference multiplies x*g by x again instead of by y).  `strict=False` gives the
    mathematically correct sum(x*y*g) for comparison.
    """

    @staticmethod
    def forward(ctx, x, y, gamma, strict):
        ctx.save_for_backward(x, y, gamma)
        ctx.strict = strict
        return (gamma.reshape(()) * y + 1) * x

    @staticmethod
    def backward(ctx, g):
        x, y, gamma = ctx.saved_tensors
        s = gamma.reshape(())
        xg = x * g
        dx = (y * s + 1) * g
        dy = xg * s
        dgamma = (xg * (x if ctx.strict else y)).sum().reshape(gamma.shape)
        return dx, dy, dgamma, None


def gate(x, y, gamma, strict_reference: bool = True):
    return _Gate.apply(x, y, gamma, strict_reference)


class _LiveSigma(torch.autograd.Function):
    """sigma = u . (W v) whose backward reads u and v AT BACKWARD TIME.

    In the reference u/v are Parameters updated through `.data = ...` (spectral_norm.py:28-29),
    which swaps storage without bumping the version counter, and autograd saved the Parameter
    objects themselves (spectral_norm.py:31).  So when the SAME SpectralNorm module is forwarded
    several times before one backward (the three discriminator passes of a D step,
    main.py:149-156), every pass's d(sigma)/dW = u v^T term is evaluated with the LATEST u, v.
    Verified against the reference's gradients in tests/golden/step_*.pt.

    u and v themselves are requires_grad=False Parameters (spectral_norm.py:45-46) -- until the training loop's
    `dis.requires_grad_(True)` (main.py:172) switches them on for the discriminator.  From then on autograd hands them
    d(sigma)/du = W v (the mv OUTPUT saved at forward time, i.e. with that pass's v) and d(sigma)/dv = W^T u (u read
    at backward time) and Nadam moves them (tests/golden/steps3_*.pt).
    """

    @staticmethod
    def forward(ctx, mat, u, v):
        ctx.u, ctx.v = u, v          # live references, deliberately not save_for_backward
        ctx.mat = mat.detach()
        ctx.wv = ctx.mat.mv(v.detach())     # torch.dot saved this tensor: fixed at forward time
        return torch.dot(u.detach(), ctx.wv)

    @staticmethod
    def backward(ctx, g):
        du = g * ctx.wv if ctx.needs_input_grad[1] else None
        dv = g * ctx.mat.t().mv(ctx.u.detach()) if ctx.needs_input_grad[2] else None
        return g * torch.outer(ctx.u.detach(), ctx.v.detach()), du, dv


def power_iterate(state: State, prefix: str, eps: float = 1e-12) -> torch.Tensor:
    """One power iteration, run on EVERY forward (spectral_norm.py:21-32,57-59).

    Mutates state[prefix+'weight_u'/'weight_v'] in place (no grad) and returns the normalised
    weight W_bar / sigma with sigma = u . (W v) differentiable through W_bar only.
    height = W_bar.shape[0] (= Cin for ConvTranspose weights, spectral_norm.py:26).
    """
    w = state[prefix + "weight_bar"]
    u = state[prefix + "weight_u"]
    v = state[prefix + "weight_v"]
    mat = w.reshape(w.shape[0], -1)
    with torch.no_grad():
        t = mat.t().mv(u)
        v.copy_(t / (t.norm() + eps))
        s = mat.mv(v)
        u.copy_(s / (s.norm() + eps))
    return w / _LiveSigma.apply(mat, u, v)


# --------------------------------------------------------------------------------------
# sub-graphs
# --------------------------------------------------------------------------------------
def _conv_pair(state: State, prefix: str, x, transpose: bool, k: int, stride: int, pad: int, growth: int, separable=False):
    """ActivatedBaseConv: conv1x1(act(convkxk(act(x)))), both spectral-normed, no bias (conv.py:11-24); `separable`:
    the k x k conv is depthwise, groups = in_features (conv.py:17)."""
    w0 = power_iterate(state, prefix + "conv_0.module.")
    h = roottanh(x, growth)
    groups = x.shape[1] if separable else 1
    if transpose:
        h = F.conv_transpose2d(h, w0, None, stride=stride, padding=pad, groups=groups)
    else:
        h = F.conv2d(h, w0, None, stride=stride, padding=pad, groups=groups)
    w1 = power_iterate(state, prefix + "conv_1.module.")
    h = roottanh(h, growth)
    if transpose:
        return F.conv_transpose2d(h, w1, None)
    return F.conv2d(h, w1, None)


def _deep_conv_plan(cfg: OracleConfig, cin: int, cout: int, transpose: bool, stride: int,
                    use_bottleneck: bool, depth: int):
    """Layer list of DeepResidualConv (conv.py:27-67): tuples
    (in, out, transpose, k, stride, pad, normalize, residual)."""
    low = min(cin, cout)
    if use_bottleneck and max(cin, cout) // low < cfg.BOTTLENECK:
        low //= cfg.BOTTLENECK
    k = 2 * stride + (0 if transpose else 1)                       # conv.py:36
    pad = max(k // 2 - stride // 2, 0) if transpose else k // 2   # utils.py:34-39
    plan = [(cin, low if depth > 1 else cout, transpose, k, stride, pad, False, False)]
    for i in range(depth - 2):
        plan.append((low, low, False, 5, 1, 2, bool(i), True))
    if depth > 1:
        plan.append((low, cout, False, 5, 1, 2, bool(depth - 2), low == cout))
    return plan


def deep_conv(state: State, prefix: str, x, cfg: OracleConfig, cin, cout, transpose, stride,
              use_bottleneck=True, depth=1, strict_reference=True):
    """DeepResidualConv.forward (conv.py:69-72)."""
    for j, (ci, co, tr, k, s, pad, normalize, residual) in enumerate(
            _deep_conv_plan(cfg, cin, cout, transpose, stride, use_bottleneck, depth)):
        name = f"{prefix}conv_{j}."
        inner = name
        if residual:
            inner += "layer_module."
        h = x
        if normalize:
            h = whole_tensor_norm(h, state[inner + "i_norm.weight"], state[inner + "i_norm.bias"])
            inner += "module."
        h = _conv_pair(state, inner, h, tr, k, s, pad, cfg.ROOTTANH_GROWTH, cfg.SEPARABLE)
        x = gate(x, h, state[name + "gamma"], strict_reference) if residual else h
    return x


def feature_pool(x, out_features: int):
    """FeaturePooling (scale.py:12-16): regroup the NCHW-contiguous memory as [B,out,H,W,r] and
    average the last axis -- i.e. r consecutive MEMORY elements, not r channels."""
    b, c, h, w = x.shape
    return x.contiguous().reshape(b, out_features, h, w, c // out_features).mean(dim=-1)


def skip_path(state: State, prefix: str, x, cin: int, cout: int, stride: int, transpose: bool):
    """Scale() (scale.py:19-45).  `prefix` is '<...>scale_layer.' (or 'residual_module.')."""
    n_layers = int(cin != cout) + int(stride > 1)
    first = prefix + ("0." if n_layers > 1 else "")
    if cin > cout:
        if cin % cout == 0:
            x = feature_pool(x, cout)
        else:
            w = power_iterate(state, first + "module.")
            x = F.conv2d(x, w, state[first + "module.bias"])
    elif cout > cin:
        w = power_iterate(state, first + "layer_module.module.")
        extra = F.conv2d(x, w, state[first + "layer_module.module.bias"])
        x = torch.cat([x, extra], dim=1)                            # merge.py:14-15
    if stride > 1:
        if transpose:
            x = F.interpolate(x, scale_factor=stride, mode="bilinear", align_corners=False)
        else:
            x = F.avg_pool2d(x, stride, stride)
    return x


def feature_attention(state: State, prefix: str, x, size: int, features: int, growth: int, separable=False, bottleneck=4):
    """attention.py:9-37: (S x 1) conv -> act -> (1 x S) conv -> act -> 1x1 conv -> softmax over channels -> broadcast
    to [B,F,S,S]; `separable` (attention.py:15-21): ONE grouped full-extent S x S conv F -> F/4 (groups = F/4, no
    activation) instead of the two axis convs."""
    if separable and features % (features // bottleneck) == 0:
        w = power_iterate(state, prefix + "0.module.")
        h = F.conv2d(x, w, groups=features // bottleneck)
        w = power_iterate(state, prefix + "1.module.")
        h = torch.softmax(F.conv2d(h, w), dim=1)
        return h.reshape(h.shape[0], -1, 1, 1).expand(-1, features, size, size)
    w = power_iterate(state, prefix + "0.module.")
    h = roottanh(F.conv2d(x, w), growth)
    w = power_iterate(state, prefix + "2.module.")
    h = roottanh(F.conv2d(h, w), growth)
    w = power_iterate(state, prefix + "4.module.")
    h = torch.softmax(F.conv2d(h, w), dim=1)
    return h.reshape(h.shape[0], -1, 1, 1).expand(-1, features, size, size)


def self_attention(state: State, prefix: str, x, growth: int):
    """attention.py:40-54: softmax over HW of W1 . act(W0 . X), X = x.view(B,F,HW)."""
    b, f = x.shape[:2]
    flat = x.reshape(b, f, -1)
    w0 = power_iterate(state, prefix + "conv_0.module.")
    h = roottanh(F.conv1d(flat, w0), growth)
    w1 = power_iterate(state, prefix + "conv_1.module.")
    h = torch.softmax(F.conv1d(h, w1), dim=-1)
    return h.reshape(x.shape)


def block(state: State, prefix: str, x, cfg: OracleConfig, out_size: int, cin: int, cout: int,
          stride: int, transpose: bool, number: int, scales=None, strict_reference=True):
    """Block.forward (block.py:44-52)."""
    scales = scales or [None, None, None]
    g = cfg.ROOTTANH_GROWTH

    def normed(t, sub, scale):
        gain = scale if scale is not None else state[prefix + sub + "layer_module.i_norm.weight"]
        return whole_tensor_norm(t, gain, state[prefix + sub + "layer_module.i_norm.bias"])

    skip = skip_path(state, prefix + "scale_layer.", x, cin, cout, stride, transpose)
    h = normed(x, "res_module_i.", scales[0])
    h = deep_conv(state, prefix + "res_module_i.layer_module.module.", h, cfg, cin, cout,
                  transpose, stride, True, cfg.DEPTH, strict_reference)
    out = gate(skip, h, state[prefix + "res_module_i.gamma"], strict_reference)
    if _has_attention(cfg, out_size, number):
        a = normed(out, "res_module_f.", scales[1])
        a = feature_attention(state, prefix + "res_module_f.layer_module.module.", a, out_size, cout, g, cfg.SEPARABLE,
                              cfg.BOTTLENECK)
        out = gate(out, a, state[prefix + "res_module_f.gamma"], strict_reference)
        a = normed(out, "res_module_s.", scales[2])
        a = self_attention(state, prefix + "res_module_s.layer_module.module.", a, g)
        out = gate(out, a, state[prefix + "res_module_s.gamma"], strict_reference)
    return out


def generator_forward(state: State, z, const_noise, cfg: OracleConfig, strict_reference=True):
    """Generator.forward (models.py:61-66) + BlockBlock.forward style chain (block.py:112-127).
    const_noise is the generator's constant [1,Z,2,2] input (`Generator.noise`, models.py:59)."""
    feats = generator_features(cfg)
    x = const_noise.expand(z.shape[0], -1, -1, -1)
    carry = None
    lin = 0
    size = 2
    for i in range(len(feats) - 1):
        size *= cfg.G_STRIDE
        n_style = 3 if _has_attention(cfg, size, i) else 1
        scales = []
        for _ in range(n_style):
            carry = z if carry is None else torch.cat([z, carry], dim=1)
            p = f"conv_block.mul_block_{lin}.module.module."
            pre = F.linear(carry, power_iterate(state, p), state[p + "bias"])   # linear.py:13-15
            carry = roottanh(pre, cfg.ROOTTANH_GROWTH)
            scales.append(pre.reshape(pre.shape[0], -1, 1, 1))
            lin += 1
        x = block(state, f"conv_block.block_{i}.", x, cfg, size, feats[i], feats[i + 1],
                  cfg.G_STRIDE, True, i, scales, strict_reference)
    x = deep_conv(state, "out_conv.", x, cfg, feats[-1], 3, False, 1, False, 1, strict_reference)
    return torch.tanh(x)


def discriminator_forward(state: State, img, cfg: OracleConfig, strict_reference=True):
    """Discriminator.forward (models.py:69-97), END_LAYER = 1."""
    feats = discriminator_features(cfg)
    skip = skip_path(state, "main.0.residual_module.", img, 3, feats[0], 2, False)
    h = deep_conv(state, "main.0.layer_module.", img, cfg, 3, feats[0], False, 2, False, 1, strict_reference)
    x = gate(skip, h, state["main.0.gamma"], strict_reference)
    size = cfg.IMAGE_SIZE // 2
    for i in range(len(feats) - 1):
        size = int(size / cfg.D_STRIDE + 1 - 1e-12)                # block.py:66-70
        x = block(state, f"main.1.block_{i}.", x, cfg, size, feats[i], feats[i + 1],
                  cfg.D_STRIDE, False, i, None, strict_reference)
    return deep_conv(state, "main.2.", x, cfg, feats[-1], 1, False, 1, False, 1, strict_reference)


# --------------------------------------------------------------------------------------
# losses, optimiser, step
# --------------------------------------------------------------------------------------
def hinge(t):  # utils.py:133-134
    return (1 - t).clamp(min=0)


def consistency_penalty(d_true, d_aug, gamma: float = 100.0):  # grad_penalty.py:1-2
    return gamma * (d_true.mean() - d_aug.reshape(-1).mean()) ** 2


class Nadam:
    """Restatement of nadam.py:31-89 over a name->tensor state (only tensors with a .grad move)."""

    def __init__(self, lr, betas=(0.9, 0.999), eps=1e-8, schedule_decay=4e-3):
        self.lr, self.betas, self.eps, self.schedule_decay = lr, betas, eps, schedule_decay
        self.slots: Dict[str, dict] = {}

    @torch.no_grad()
    def step(self, state: State):
        b1, b2 = self.betas
        for name, p in state.items():
            if p.grad is None:
                continue
            g = p.grad
            s = self.slots.setdefault(name, {"t": 0, "m_sched": 1.0,
                                             "m": torch.zeros_like(p), "v": torch.zeros_like(p)})
            s["t"] += 1
            t = s["t"]
            mu_t = b1 * (1.0 - 0.5 * 0.96 ** (t * self.schedule_decay))
            mu_n = b1 * (1.0 - 0.5 * 0.96 ** ((t + 1) * self.schedule_decay))
            sched_new = s["m_sched"] * mu_t
            sched_next = sched_new * mu_n
            s["m_sched"] = sched_new
            s["m"].mul_(b1).add_(g, alpha=1.0 - b1)
            s["v"].mul_(b2).addcmul_(g, g, value=1.0 - b2)
            denom = (s["v"] / (1.0 - b2 ** t)).sqrt_().add_(self.eps)
            p.addcdiv_(g, denom, value=-self.lr * (1.0 - mu_t) / (1.0 - sched_new))
            p.addcdiv_(s["m"], denom, value=-self.lr * mu_n / (1.0 - sched_next))


def _zero_grads(state: State):
    for p in state.values():
        p.grad = None


def train_step(g_state: State, d_state: State, const_noise, real, aug, z, cfg: OracleConfig,
               g_opt: Optional[Nadam] = None, d_opt: Optional[Nadam] = None,
               strict_reference=True) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """One G+D step = main.py:142-172 with miniter = MINIBATCHES = DITERS = 1.
    Returns (d_hinge_loss, penalty, g_loss); leaves gradients in .grad of both states."""
    with torch.no_grad():
        fake = generator_forward(g_state, z, const_noise, cfg, strict_reference)
    _zero_grads(d_state)
    d_true = discriminator_forward(d_state, real, cfg, strict_reference).reshape(-1)
    d_gen = -discriminator_forward(d_state, fake, cfg, strict_reference).reshape(-1)
    d_err = (hinge(d_true) + hinge(d_gen)).mean()
    pen = consistency_penalty(d_true, discriminator_forward(d_state, aug, cfg, strict_reference))
    (d_err + pen).backward()
    if d_opt is not None:
        d_opt.step(d_state)
    # main.py:160,172: dis.requires_grad_(False) ... dis.requires_grad_(True) -- the second call also switches on the
    # discriminator's weight_u / weight_v (see _LiveSigma); the non-strict mode keeps them frozen
    trainable = {k: (True if strict_reference else p.requires_grad) for k, p in d_state.items()}
    for p in d_state.values():
        p.requires_grad_(False)
    _zero_grads(g_state)
    g_err = hinge(discriminator_forward(d_state, generator_forward(g_state, z, const_noise, cfg, strict_reference),
                                        cfg, strict_reference).reshape(-1)).mean()
    g_err.backward()
    if g_opt is not None:
        g_opt.step(g_state)
    for k, p in d_state.items():
        p.requires_grad_(trainable[k])
    return d_err.detach(), pen.detach(), g_err.detach()


# --------------------------------------------------------------------------------------
# state construction (distributions of the reference constructors + utils.init)
# --------------------------------------------------------------------------------------
def _sn_entries(state: State, prefix: str, shape, bias: bool, gen, dtype):
    fan_in = int(torch.tensor(shape[1:]).prod()) if len(shape) > 1 else shape[0]
    bound = 1.0 / math.sqrt(fan_in)                                # kaiming_uniform(a=sqrt(5))
    w = (torch.rand(shape, generator=gen, dtype=dtype) * 2 - 1) * bound
    height, width = shape[0], w.reshape(shape[0], -1).shape[1]
    u = torch.randn(height, generator=gen, dtype=dtype)
    v = torch.randn(width, generator=gen, dtype=dtype)
    if bias:
        n_bias = shape[0]
        state[prefix + "bias"] = torch.zeros(n_bias, dtype=dtype)  # utils.py:127-130
    state[prefix + "weight_u"] = u / (u.norm() + 1e-12)
    state[prefix + "weight_v"] = v / (v.norm() + 1e-12)
    state[prefix + "weight_bar"] = w


def _gamma(gen, m: int, dtype):
    sign = 1.0 if torch.rand((), generator=gen).item() < 0.5 else -1.0   # orthogonal_ of a 1x1
    return torch.full((1, 1), sign + m + 1, dtype=dtype)                 # merge.py:51-53


def _norm_entries(state, prefix, c, gen, dtype):
    state[prefix + "weight"] = 0.998 + 0.004 * torch.rand((1, c, 1, 1), generator=gen, dtype=dtype)
    state[prefix + "bias"] = torch.zeros((1, c, 1, 1), dtype=dtype)


def _deep_conv_entries(state, prefix, cfg, cin, cout, transpose, stride, use_bottleneck, depth, gen, dtype):
    for j, (ci, co, tr, k, s, pad, normalize, residual) in enumerate(
            _deep_conv_plan(cfg, cin, cout, transpose, stride, use_bottleneck, depth)):
        name = f"{prefix}conv_{j}."
        inner = name + ("layer_module." if residual else "")
        if residual:
            state[name + "gamma"] = _gamma(gen, 1, dtype)
        if normalize:
            _norm_entries(state, inner + "i_norm.", ci, gen, dtype)
            inner += "module."
        mid = 1 if cfg.SEPARABLE else ci               # depthwise: one filter per channel (conv.py:17)
        if tr:
            _sn_entries(state, inner + "conv_0.module.", (ci, mid, k, k), False, gen, dtype)
            _sn_entries(state, inner + "conv_1.module.", (ci, co, 1, 1), False, gen, dtype)
        else:
            _sn_entries(state, inner + "conv_0.module.", (ci, mid, k, k), False, gen, dtype)
            _sn_entries(state, inner + "conv_1.module.", (co, ci, 1, 1), False, gen, dtype)


def _block_entries(state, prefix, cfg, out_size, cin, cout, stride, transpose, number, gen, dtype):
    n_layers = int(cin != cout) + int(stride > 1)
    first = prefix + "scale_layer." + ("0." if n_layers > 1 else "")
    if cin > cout and cin % cout != 0:
        _sn_entries(state, first + "module.", (cout, cin, 1, 1), True, gen, dtype)
    elif cout > cin:
        _sn_entries(state, first + "layer_module.module.", (cout - cin, cin, 1, 1), True, gen, dtype)
    state[prefix + "res_module_i.gamma"] = _gamma(gen, 3, dtype)
    _norm_entries(state, prefix + "res_module_i.layer_module.i_norm.", cin, gen, dtype)
    _deep_conv_entries(state, prefix + "res_module_i.layer_module.module.", cfg, cin, cout, transpose,
                       stride, True, cfg.DEPTH, gen, dtype)
    if _has_attention(cfg, out_size, number):
        bf = cout // cfg.BOTTLENECK
        state[prefix + "res_module_f.gamma"] = _gamma(gen, 0, dtype)
        _norm_entries(state, prefix + "res_module_f.layer_module.i_norm.", cout, gen, dtype)
        fa = prefix + "res_module_f.layer_module.module."
        if cfg.SEPARABLE and cout % bf == 0:               # attention.py:15-21
            _sn_entries(state, fa + "0.module.", (bf, cout // bf, out_size, out_size), False, gen, dtype)
            _sn_entries(state, fa + "1.module.", (cout, bf, 1, 1), False, gen, dtype)
        else:
            _sn_entries(state, fa + "0.module.", (bf, cout, out_size, 1), False, gen, dtype)
            _sn_entries(state, fa + "2.module.", (bf, bf, 1, out_size), False, gen, dtype)
            _sn_entries(state, fa + "4.module.", (cout, bf, 1, 1), False, gen, dtype)
        state[prefix + "res_module_s.gamma"] = _gamma(gen, 0, dtype)
        _norm_entries(state, prefix + "res_module_s.layer_module.i_norm.", cout, gen, dtype)
        sa = prefix + "res_module_s.layer_module.module."
        _sn_entries(state, sa + "conv_0.module.", (cout, cout, 1), False, gen, dtype)
        _sn_entries(state, sa + "conv_1.module.", (cout, cout, 1), False, gen, dtype)


def init_generator_state(cfg: OracleConfig, seed: int = 999, dtype=torch.float32) -> Tuple[State, torch.Tensor]:
    """Random generator state with the reference's names, shapes and init distributions
    (NOT its RNG stream).  Returns (state, const_noise)."""
    gen = torch.Generator().manual_seed(seed)
    feats = generator_features(cfg)
    z = cfg.INPUT_VECTOR_Z
    state: State = {}
    size, prev_out, lin = 2, 0, 0
    for i in range(len(feats) - 1):
        size *= cfg.G_STRIDE
        cin, cout = feats[i], feats[i + 1]
        _block_entries(state, f"conv_block.block_{i}.", cfg, size, cin, cout, cfg.G_STRIDE, True, i, gen, dtype)
    for i in range(len(feats) - 1):                                  # block.py:80-101
        cin, cout = feats[i], feats[i + 1]
        attn = _has_attention(cfg, 2 * cfg.G_STRIDE ** (i + 1), i)
        group_in = prev_out if (prev_out and prev_out != cin) else cin
        dims = [(group_in + z * bool(i), cin)]
        if attn:
            dims += [(cin + z, cout), (cout + z, cout)]
            prev_out = cout
        else:
            prev_out = cin
        for fin, fout in dims:
            _sn_entries(state, f"conv_block.mul_block_{lin}.module.module.", (fout, fin), True, gen, dtype)
            lin += 1
    _deep_conv_entries(state, "out_conv.", cfg, feats[-1], 3, False, 1, False, 1, gen, dtype)
    const_noise = torch.randn((1, z, 2, 2), generator=gen, dtype=dtype)
    return _as_leaves(state), const_noise


def init_discriminator_state(cfg: OracleConfig, seed: int = 1000, dtype=torch.float32) -> State:
    gen = torch.Generator().manual_seed(seed)
    feats = discriminator_features(cfg)
    state: State = {"main.0.gamma": _gamma(gen, 0, dtype)}
    _sn_entries(state, "main.0.residual_module.0.layer_module.module.", (feats[0] - 3, 3, 1, 1), True, gen, dtype)
    _deep_conv_entries(state, "main.0.layer_module.", cfg, 3, feats[0], False, 2, False, 1, gen, dtype)
    size = cfg.IMAGE_SIZE // 2
    for i in range(len(feats) - 1):
        size = int(size / cfg.D_STRIDE + 1 - 1e-12)
        _block_entries(state, f"main.1.block_{i}.", cfg, size, feats[i], feats[i + 1], cfg.D_STRIDE, False, i, gen, dtype)
    _deep_conv_entries(state, "main.2.", cfg, feats[-1], 1, False, 1, False, 1, gen, dtype)
    return _as_leaves(state)


def _as_leaves(state: State) -> State:
    """u/v are requires_grad=False parameters in the reference (spectral_norm.py:45-46)."""
    out = {}
    for k, t in state.items():
        t = t.detach().clone()
        t.requires_grad_(not (k.endswith("weight_u") or k.endswith("weight_v")))
        out[k] = t
    return out


def load_state(tensors: Dict[str, torch.Tensor], dtype=None) -> State:
    """Wrap a reference state_dict (or golden fixture) as oracle leaves."""
    return _as_leaves({k: (v.to(dtype) if dtype is not None else v) for k, v in tensors.items()})


def synthetic_batch(cfg: OracleConfig, batch: int, dtype=torch.float32):
    """SURVEY.md section 8d synthetic inputs: real (seed 0), aug (seed 1), z (seed 2)."""
    s = cfg.IMAGE_SIZE
    real = torch.randn((batch, 3, s, s), generator=torch.Generator().manual_seed(0)).clamp(-1, 1)
    aug = (real + 0.05 * torch.randn((batch, 3, s, s), generator=torch.Generator().manual_seed(1))).clamp(-1, 1)
    z = torch.randn((batch, cfg.INPUT_VECTOR_Z), generator=torch.Generator().manual_seed(2))
    return real.to(dtype), aug.to(dtype), z.to(dtype)
