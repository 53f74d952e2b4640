"""The reference's nn.Module API rebuilt on the sm_100a kernels.

Same class names, constructor signatures, forward signatures, attribute names and state_dict keys
as /root/reference/libs (activation.py, inplace_norm.py, spectral_norm.py, merge.py, scale.py,
attention.py, conv.py, linear.py, block.py), so the reference's models.py / main.py can import these
instead.  The arithmetic lives in locate_b200.ops (kernels); nothing here computes with torch ops
on the data path.  Parameters are created by the same torch constructors in the same order as the
reference, so `torch.manual_seed(s)` yields the reference's initial weights.
"""
import torch
from torch import nn

from . import ops
from .config import CFG
from .ops import ConvSpec


def identity(x):
    return x


# ---- activation (libs/activation.py:39-51) ---------------------------------------------------
def nonlinear_function(x):
    return ops.roottanh(x, CFG.ROOTTANH_GROWTH)


class RootTanhModule(nn.Module):
    @staticmethod
    def forward(function_input):
        return nonlinear_function(function_input)


NonLinear = RootTanhModule


# ---- whole-tensor norm (libs/inplace_norm.py:34-56) --------------------------------------------
class InPlaceNorm(nn.Module):
    def __init__(self, features=1, dim=2):
        super().__init__()
        self.weight = nn.Parameter(torch.ones((1, features, *[1] * dim)))
        self.bias = nn.Parameter(torch.zeros((1, features, *[1] * dim)))

    def forward(self, function_input, scale=None, emit=None, _mailbox=None):
        return ops.whole_norm(function_input, self.weight if scale is None else scale, self.bias, emit, _mailbox)


def _emit_mode(module):
    """Which bf16 GEMM operand the norm kernel should leave for the module that follows it: "act" when the module starts
    with RootTanh -> conv (conv.py:22-24), "plain" when a spectral-normed conv reads the norm output directly
    (attention.py:26,44) and nothing else does, None otherwise."""
    if isinstance(module, ActivatedBaseConv):
        return None if module.conv_0.spec.groups > 1 else "act"     # the depthwise kernel applies RootTanh itself
    if isinstance(module, DeepResidualConv):
        return _emit_mode(module.layers[0]) if len(module.layers) == 1 else _emit_mode_first(module.layers[0])
    if isinstance(module, SelfAttention):
        return "plain"
    if isinstance(module, nn.Sequential) and len(module) > 0 and isinstance(module[0], SpectralNorm):
        return "plain"
    return None


def _emit_mode_first(layer):
    return "act" if isinstance(layer, ActivatedBaseConv) else None


class Norm(nn.Module):
    def __init__(self, features, module, dim=2):
        super().__init__()
        self.i_norm = InPlaceNorm(features, dim=dim)
        self.module = module

    def forward(self, function_input, scale=None, _mailbox=None):
        return self.module(self.i_norm(function_input, scale, emit=_emit_mode(self.module), _mailbox=_mailbox))


# ---- spectral norm (libs/spectral_norm.py:12-59) -----------------------------------------------
def _pair(v):
    return (v, v) if isinstance(v, int) else tuple(v)


def _spec_of(module):
    """ConvSpec of the torch layer the reference wraps in SpectralNorm."""
    if isinstance(module, nn.Linear):
        return ConvSpec("linear", module.in_features, module.out_features)
    if isinstance(module, nn.Conv1d):
        if module.kernel_size != (1,) or module.stride != (1,) or module.padding != (0,) or module.groups != 1:
            raise NotImplementedError("only Conv1d(kernel_size=1) is on the reference's path (attention.py:44-46)")
        return ConvSpec("conv1d", module.in_channels, module.out_channels)
    if isinstance(module, (nn.Conv2d, nn.ConvTranspose2d)):
        transposed = isinstance(module, nn.ConvTranspose2d)
        (kh, kw), (sh, sw), (ph, pw) = module.kernel_size, module.stride, _pair(module.padding)
        if module.dilation != (1, 1) or sh != sw or ph != pw or module.padding_mode != "zeros":
            raise NotImplementedError("dilation / anisotropic stride or padding are not on the reference's path")
        if transposed and module.output_padding != (0, 0):
            raise NotImplementedError("output_padding")
        g = module.groups
        if g != 1:
            # SEPARABLE = True builds exactly two grouped forms: depthwise (conv.py:17) and the full-extent
            # feature-attention conv with one output per group (attention.py:15-21)
            depthwise = g == module.in_channels == module.out_channels
            full = (not transposed) and g == module.out_channels and module.in_channels % g == 0 and (sh, ph) == (1, 0)
            if not (depthwise or full):
                raise NotImplementedError(f"grouped convolution groups={g} {module.in_channels}->{module.out_channels}")
        return ConvSpec("convT" if transposed else "conv", module.in_channels, module.out_channels, kh, kw, sh, ph, g)
    raise NotImplementedError(f"SpectralNorm over {type(module).__name__}")


class SpectralNorm(nn.Module):
    """Wraps a torch Conv2d / ConvTranspose2d / Conv1d(k=1) / Linear exactly like the reference:
    `weight` is replaced by `weight_u`, `weight_v` (requires_grad=False) and `weight_bar`; every
    forward runs one power iteration and applies the layer with weight_bar / sigma -- here as one
    fused kernel sequence (1/sigma lives in the GEMM epilogue; weight_bar is never rescaled)."""

    def __init__(self, module, name='weight', power_iterations=1):
        super().__init__()
        self.module = module
        self.name = name
        self.power_iterations = power_iterations
        if name != 'weight':
            raise NotImplementedError("the reference only ever normalises `weight`")
        self.spec = _spec_of(module)
        self._pre_sigma = None            # set by sn_batch.SpectralBatch.run() when the model iterates all layers at once
        if not self._made_params():
            self._make_params()

    def _made_params(self):
        return all(hasattr(self.module, self.name + s) for s in ("_u", "_v", "_bar"))

    def _make_params(self):
        w = getattr(self.module, self.name)
        height = w.data.shape[0]
        width = w.data.numel() // height
        u = nn.Parameter(w.data.new(height).normal_(0, 1), requires_grad=False)
        v = nn.Parameter(w.data.new(width).normal_(0, 1), requires_grad=False)
        u.data = u.data / (u.data.norm() + 1e-12)
        v.data = v.data / (v.data.norm() + 1e-12)
        w_bar = nn.Parameter(w.data)
        del self.module._parameters[self.name]
        self.module.register_parameter(self.name + "_u", u)
        self.module.register_parameter(self.name + "_v", v)
        self.module.register_parameter(self.name + "_bar", w_bar)

    def forward(self, *args, cat_input=False, pre_act=False):
        """pre_act=True applies RootTanh to the input inside the same kernel sequence (the activation that
        precedes both convs of ActivatedBaseConv, conv.py:23-24)."""
        (x,) = args
        m = self.module
        for _ in range(self.power_iterations - 1):        # extra iterations only move u/v
            ops.power_iterate(m.weight_bar, m.weight_u.data, m.weight_v.data, self.spec)
        squeeze = x.dim() == 3                            # Conv1d input [B,F,L]
        if squeeze:
            x = x.unsqueeze(-1)
        pre_sigma, self._pre_sigma = self._pre_sigma, None
        out = ops.sn_conv(x, m.weight_bar, m.weight_u, m.weight_v, getattr(m, "bias", None), self.spec, cat_input, pre_act,
                          pre_sigma)
        return out.squeeze(-1) if squeeze else out


# ---- merge (libs/merge.py:4-62) ----------------------------------------------------------------
class CatModule(nn.Module):
    def __init__(self, residual_module, layer_module):
        super().__init__()
        self.residual_module = residual_module
        self.layer_module = layer_module

    def forward(self, function_input, layer_input=None, scale=None):
        args = [function_input] if layer_input is None else [layer_input]
        if scale is not None:
            args.append(scale)
        if (self.residual_module is identity and isinstance(self.layer_module, SpectralNorm)
                and len(args) == 1 and args[0] is function_input and function_input.dim() == 4):
            return self.layer_module(function_input, cat_input=True)   # GEMM writes into the concat slice
        res, layer_out = self.residual_module(function_input), self.layer_module(*args)
        return ops.CatFn.apply(res, layer_out)


def residual_function(x, attention, gamma, _mailbox=None):
    return ops.gate(x, attention, gamma, CFG.STRICT_REFERENCE, _mailbox)


class ResModule(nn.Module):
    def __init__(self, residual_module, layer_module, m=0):
        super().__init__()
        self.residual_module = residual_module
        self.layer_module = layer_module
        self.gamma = nn.Parameter(torch.ones((1, 1)))
        nn.init.orthogonal_(self.gamma.data)
        self.gamma.data.add_(m + 1)

    def _broadcast_tail(self):
        """layer_module = Norm(Sequential[..., Expand]) (feature attention): the Expand is folded into
        the gate kernel instead of materialising [B,F,S,S]."""
        lm = self.layer_module
        return (isinstance(lm, Norm) and isinstance(lm.module, nn.Sequential) and len(lm.module) > 0
                and isinstance(lm.module[-1], Expand))

    def forward(self, function_input, layer_input=None, scale=None):
        args = [function_input] if layer_input is None else [layer_input]
        if scale is not None:
            args.append(scale)
        res = self.residual_module(function_input)
        # the gate and the norm read the SAME tensor: their two gradients meet in the norm's backward kernel (ops.GradMailbox)
        mb = None
        if res is function_input and layer_input is None and isinstance(self.layer_module, Norm) and torch.is_grad_enabled():
            mb = ops.GradMailbox()
        if self._broadcast_tail():
            h = self.layer_module.i_norm(*args, emit=_emit_mode(self.layer_module.module), _mailbox=mb)
            for layer in list(self.layer_module.module)[:-1]:
                h = layer(h)
            layer_out = h                                  # [B,F,1,1] gate
        elif mb is not None:
            layer_out = self.layer_module(*args, _mailbox=mb)
        else:
            layer_out = self.layer_module(*args)
        return residual_function(res, layer_out, self.gamma, mb)


# ---- helpers (libs/util_modules.py:6-12, libs/utils.py:34-46) ----------------------------------
class Expand(nn.Module):
    def __init__(self, *target_size):
        super().__init__()
        self.target_size = target_size

    def forward(self, function_input):
        t = function_input
        return t.view(t.size(0), -1, *[1] * (len(t.size()) - 2)).expand(self.target_size)


class SoftmaxChannels(nn.Module):
    """torch.nn.Softmax(dim=1) of attention.py:35."""

    @staticmethod
    def forward(function_input):
        return ops.SoftmaxChannelsFn.apply(function_input)


def conv_pad_tuple(kernel_size, _, dim=2):
    return tuple([kernel_size // 2] * dim)


def transpose_pad_tuple(kernel_size, stride, dim=2):
    return tuple([max(kernel_size // 2 - stride // 2, 0)] * dim)


def _need_2d(dim):
    if dim != 2:
        raise NotImplementedError("the reference's models are 2-D (dim=2)")


# ---- scale / skip path (libs/scale.py:7-45) ----------------------------------------------------
class FeaturePooling(nn.Module):
    def __init__(self, out_features):
        super().__init__()
        self.out_features = out_features

    def forward(self, function_input):
        return ops.FeaturePoolFn.apply(function_input, self.out_features)


class BilinearUp2(nn.Module):
    """nn.Upsample(mode='bilinear', scale_factor=2, align_corners=False)."""

    @staticmethod
    def forward(function_input):
        return ops.Upsample2xFn.apply(function_input)


class AvgPool2(nn.Module):
    """nn.AvgPool2d(2, 2)."""

    @staticmethod
    def forward(function_input):
        return ops.AvgPool2Fn.apply(function_input)


def Scale(in_features, out_features, stride, transpose, dim=2):
    _need_2d(dim)
    reslayers = []
    if in_features > out_features:
        if in_features % out_features == 0:
            reslayers.append(FeaturePooling(out_features))
        else:
            reslayers.append(SpectralNorm(nn.Conv2d(in_features, out_features, 1)))
    elif out_features > in_features:
        reslayers.append(CatModule(identity, SpectralNorm(nn.Conv2d(in_features, out_features - in_features, 1))))
    if stride > 1:
        if stride != 2:
            raise NotImplementedError("only stride 2 resampling is on the reference's path (config.py:48-49)")
        reslayers.append(BilinearUp2() if transpose else AvgPool2())
    if len(reslayers) > 1:
        return nn.Sequential(*reslayers)
    if not reslayers:
        return identity
    return reslayers[0]


# ---- attention (libs/attention.py:9-54) --------------------------------------------------------
def feature_attention(in_size, features, dim=2):
    _need_2d(dim)
    bfeatures = features // CFG.BOTTLENECK
    layers = []
    input_features = features
    min_features = min(input_features, bfeatures)
    if CFG.SEPARABLE and input_features % min_features == 0 and bfeatures % min_features == 0:
        # one grouped conv spanning the whole map, no activation after it (attention.py:15-21)
        layers.append(SpectralNorm(nn.Conv2d(input_features, bfeatures, kernel_size=[in_size] * dim, bias=False,
                                             groups=min_features)))
    else:
        for i in range(dim):
            kernel_size = [1] * dim
            kernel_size[i] = in_size
            layers.extend([SpectralNorm(nn.Conv2d(input_features, bfeatures, kernel_size=kernel_size, bias=False, groups=1)),
                           NonLinear()])
            input_features = bfeatures
    layers.extend([SpectralNorm(nn.Conv2d(bfeatures, features, kernel_size=1, bias=False)),
                   SoftmaxChannels(),
                   Expand(-1, features, *([in_size] * dim))])
    return nn.Sequential(*layers)


class SoftmaxPixels(nn.Module):
    """torch.nn.Softmax(dim=-1) on the [B,F,HW] view (attention.py:47), applied to [B,F,H,W] directly."""

    @staticmethod
    def forward(function_input):
        return ops.SoftmaxPixelsFn.apply(function_input)


class SelfAttention(nn.Module):
    def __init__(self, features):
        super().__init__()
        args = [features, features, 1]
        self.conv_0 = SpectralNorm(nn.Conv1d(*args, bias=False))
        self.nlin_0 = NonLinear()
        self.conv_1 = SpectralNorm(nn.Conv1d(*args, bias=False))
        self.nlin_1 = SoftmaxPixels()

    def forward(self, function_input):
        batch, features, *size = function_input.size()
        x = function_input
        if len(size) != 2:                                  # generic [B,F,*]: one pixel axis
            x = function_input.reshape(batch, features, -1, 1)
        fused = ops.activated_pair(x, self.conv_0, self.conv_1, pre_act0=False) if len(size) == 2 else None
        if fused is not None:                             # conv_1(RootTanh(conv_0(x))) as one node (attention.py:48-52)
            out = self.nlin_1(fused)
        else:
            out = self.nlin_1(self.conv_1(self.nlin_0(self.conv_0(x))))   # Conv1d(k=1) == per-pixel GEMM
        return out if len(size) == 2 else out.reshape(batch, features, *size)


# ---- convolutions (libs/conv.py:11-72) ---------------------------------------------------------
class ActivatedBaseConv(nn.Module):
    def __init__(self, in_features, out_features, conv, kernel=5, stride=1, pad=2):
        super().__init__()
        mid = in_features * CFG.FEATURE_MULTIPLIER
        self.conv_0 = SpectralNorm(conv(in_channels=in_features, kernel_size=kernel, stride=stride, padding=pad,
                                        bias=False, out_channels=mid, groups=in_features if CFG.SEPARABLE else 1))
        self.conv_1 = SpectralNorm(conv(kernel_size=1, stride=1, padding=0, out_channels=out_features, bias=False,
                                        in_channels=mid))

    def forward(self, function_input):
        fused = ops.activated_pair(function_input, self.conv_0, self.conv_1)
        if fused is not None:
            return fused
        return self.conv_1(self.conv_0(function_input, pre_act=True), pre_act=True)


class DeepResidualConv(nn.Module):
    def __init__(self, in_features, out_features, transpose, stride, use_bottleneck=True, dim=2, depth=1):
        super().__init__()
        _need_2d(dim)
        min_features = min(in_features, out_features)
        if use_bottleneck and max(in_features, out_features) // min_features < CFG.BOTTLENECK:
            min_features //= CFG.BOTTLENECK
        self.final_layer = None
        kernel = stride * 2 + int(not transpose)
        self.layers = []

        def add_conv(cin, cout, residual=True, normalize=False, transposed=False, conv_stride=1, **kwargs):
            conv = nn.ConvTranspose2d if transposed else nn.Conv2d
            layer = ActivatedBaseConv(cin, cout, conv, stride=conv_stride, **kwargs)
            if normalize:
                layer = Norm(cin, layer, dim)
            if residual and cin == cout:
                layer = ResModule(identity, layer, m=1)
            setattr(self, f'conv_{len(self.layers)}', layer)
            self.layers.append(layer)

        pad_tuple = transpose_pad_tuple if transpose else conv_pad_tuple
        add_conv(in_features, min_features if depth > 1 else out_features, False, False, transpose, stride,
                 kernel=kernel, pad=pad_tuple(kernel, stride))
        for i in range(depth - 2):
            add_conv(min_features, min_features, True, normalize=bool(i))
        if depth > 1:
            add_conv(min_features, out_features, True, normalize=bool(depth - 2))

    def forward(self, function_input):
        for layer in self.layers:
            function_input = layer(function_input)
        return function_input


# ---- style linear (libs/linear.py:7-15) --------------------------------------------------------
class LinearModule(nn.Module):
    def __init__(self, *args):
        super().__init__()
        self.module = SpectralNorm(nn.Linear(*args))
        self.nlin = NonLinear()

    def forward(self, function_input):
        out = self.module(function_input)
        return self.nlin(out), out


# ---- blocks (libs/block.py:15-127) -------------------------------------------------------------
class Block(nn.Module):
    def __init__(self, in_size, in_features, out_features, stride, transpose, block_number, cat_out=True, dim=2):
        super().__init__()
        self.scale_layer = Scale(in_features, out_features, stride, transpose, dim=dim)
        self.res_module_i = ResModule(identity,
                                      Norm(in_features,
                                           DeepResidualConv(in_features, out_features, transpose, stride,
                                                            depth=CFG.DEPTH, dim=dim),
                                           dim=dim),
                                      m=3)
        if in_size >= CFG.MIN_ATTENTION_SIZE and block_number % CFG.ATTENTION_EVERY_NTH_LAYER == 0:
            self.res_module_f = ResModule(identity, Norm(out_features, feature_attention(in_size, out_features, dim=dim),
                                                         dim=dim))
            self.res_module_s = ResModule(identity, Norm(out_features, SelfAttention(out_features), dim=dim))
            self.attention = True
        else:
            self.attention = False
        self.cat_out = cat_out

    def forward(self, function_input, scales=None):
        if scales is None:
            scales = [None] * 4
        scaled = self.scale_layer(function_input)
        out = self.res_module_i(scaled, function_input, scales[0])
        if self.attention:
            out = self.res_module_f(out, scale=scales[1])
            out = self.res_module_s(out, scale=scales[2])
        return out


class BlockBlock(nn.Module):
    def __init__(self, block_count, in_size, features, strides, transpose, mul_channel=False, dim=2):
        super().__init__()
        self.block_count = block_count
        factors = strides if transpose else [1 / s for s in strides]

        def feature_tuple(idx):
            return features[idx], features[idx + 1]

        def size(idx):
            out = in_size
            for f in factors[:idx + 1]:
                out = out * f
            return int(out + 1 - 1e-12)

        blocks = [Block(size(i), *feature_tuple(i), strides[i], transpose, i, dim=dim) for i in range(block_count)]
        self.blocks = blocks
        for i, block in enumerate(blocks):
            setattr(self, f'block_{i}', block)
        sums = [0]
        depths = []
        if mul_channel:
            z = CFG.INPUT_VECTOR_Z
            mul_blocks = []
            prev_out = 0
            for i in range(block_count):
                scales = 2 * blocks[i].attention
                depths.append(1 + scales)
                sums.append(sums[-1] + scales + 1)
                inp, out = feature_tuple(i)
                group_inp = prev_out if (prev_out and prev_out != inp) else inp
                mul_blocks.append(LinearModule(group_inp + z * bool(i), inp))
                if scales:
                    mul_blocks.append(LinearModule(inp + z, out))
                    mul_blocks.extend(LinearModule(out + z, out) for _ in range(1, scales))
                    prev_out = out
                else:
                    prev_out = inp
            self.mul_blocks = mul_blocks
            for i, block in enumerate(mul_blocks):
                setattr(self, f'mul_block_{i}', block)
        self.depths = depths
        self.sums = sums
        self.out_features = feature_tuple(block_count - 1)[1]

    def forward(self, function_input, noise=None):
        next_input = None
        for i in range(self.block_count):
            operand = None
            if noise is not None:
                operand = []
                for idx in range(self.depths[i]):
                    next_input = noise if next_input is None else ops.CatFn.apply(noise, next_input)
                    next_input, factor = self.mul_blocks[self.sums[i] + idx](next_input)
                    operand.append(factor.view(*factor.size(), 1, 1))
            function_input = self.blocks[i](function_input, scales=operand)
        return function_input
