"""Generator / Discriminator with the reference's constructors and forward signatures
(libs/models.py:12-97).  Pure graph assembly over locate_b200.layers; state_dict keys equal the
reference's, so its checkpoints (netG.torch / netD.torch, main.py:235-236) load unchanged."""
import torch
from torch import nn

from . import ops
from .config import CFG
from .layers import BlockBlock, DeepResidualConv, ResModule, Scale, identity


def quadnorm(number: int):
    return number // 4 * 4


def generator_feature_list():
    """[in, f(L-2), ..., f(0)] with f(i) = quadnorm(GEN_FEATURES * FACTOR**(i - (L-1)))  (models.py:16-22,37-52); in = Z, or
    with an input block (START_LAYER >= 1) its output width f(L-1) = quadnorm(GEN_FEATURES) (models.py:44-50)."""
    clayers = CFG.LAYERS - 1
    feats = [quadnorm(int(CFG.GEN_FEATURES * CFG.FACTOR ** (i - clayers))) for i in range(clayers - 1, -1, -1)]
    first = quadnorm(int(CFG.GEN_FEATURES)) if CFG.START_LAYER >= 1 else CFG.INPUT_VECTOR_Z
    return [first] + feats


def discriminator_feature_list():
    """[d(0), ..., d(L-2), d(L-2)] with d(i) = quadnorm(DIS_FEATURES * FACTOR**(i + 1 - (L-1)))  (models.py:25-31,72-78)."""
    n = CFG.LAYERS - 1
    feats = [quadnorm(int(CFG.DIS_FEATURES * CFG.FACTOR ** ((i + 1) - n))) for i in range(n)]
    return feats + [feats[-1]]


class _ArenaModel(nn.Module):
    """zero_grad() keeps the optimizer's flat gradient arena attached (one fill kernel) instead of
    dropping every .grad to None."""

    _lb_optimizer = None

    def _iterate_spectral_norms(self):
        """One power iteration for every spectral-normed layer, batched (sn_batch.py)."""
        batch = self.__dict__.get("_lb_sn_batch")
        if batch is None:
            from .sn_batch import SpectralBatch
            batch = SpectralBatch(self)
            self.__dict__["_lb_sn_batch"] = batch
        batch.run()

    def _finish_uv_grads(self):
        """Gradients of trainable spectral-norm weight_v's (the reference's main.py:172 quirk, sn_batch.py)."""
        batch = self.__dict__.get("_lb_sn_batch")
        if batch is not None:
            batch.finish_uv_grads()

    def zero_grad(self, set_to_none=True):
        opt = self.__dict__.get("_lb_optimizer")
        if opt is not None:
            opt.zero_grad()
        else:
            super().zero_grad(set_to_none=set_to_none)


class Generator(_ArenaModel):
    def __init__(self):
        super().__init__()
        if CFG.G_STRIDE != 2:
            raise NotImplementedError("G_STRIDE != 2")
        strides = [2] * (CFG.LAYERS - 1)
        feature_list = generator_feature_list()
        if CFG.START_LAYER >= 1:
            # models.py:46-48.  NOTE: the reference's own forward then fails -- the style chain's first Linear is built
            # for `feature_list[0]` inputs (block.py:88-96) but receives the Z-dimensional latent -- and so does this one
            # (same constructors, same state_dict keys, same shape error); see tests/test_abi_and_host.py.
            self.input_block = DeepResidualConv(CFG.INPUT_VECTOR_Z, feature_list[0], False, 1, False, 2, CFG.START_LAYER)
        else:
            self.input_block = identity
        self.conv_block = BlockBlock(len(strides), 2, feature_list, strides, True, True)
        self.out_conv = DeepResidualConv(self.conv_block.out_features, 3, False, 1, False, 2, 1)
        self.g_in = feature_list[0]
        self.noise = torch.randn(1, CFG.INPUT_VECTOR_Z, 2, 2)      # plain attribute, not in state_dict (models.py:59)

    def _apply(self, fn, *args, **kwargs):
        super()._apply(fn, *args, **kwargs)
        self.noise = fn(self.noise)
        return self

    def forward(self, function_input):
        self._iterate_spectral_norms()
        expanded_noise = self.noise.expand(function_input.size(0), -1, -1, -1)
        conv_out = self.input_block(expanded_noise)
        conv_out = self.conv_block(conv_out, function_input)
        conv_out = self.out_conv(conv_out)
        return ops.TanhFn.apply(conv_out)


class Discriminator(_ArenaModel):
    def __init__(self):
        super().__init__()
        if CFG.END_LAYER != 1:
            raise NotImplementedError("END_LAYER != 1")
        if CFG.D_STRIDE != 2:
            raise NotImplementedError("D_STRIDE != 2")
        strides = [2] * (CFG.LAYERS - 1)
        feature_list = discriminator_feature_list()
        first = feature_list[0]
        cat_module = ResModule(Scale(3, first, 2, False), DeepResidualConv(3, first, False, 2, False, 2, 1))
        block_block = BlockBlock(len(strides), CFG.IMAGE_SIZE // 2, feature_list, strides, False)
        tail = DeepResidualConv(block_block.out_features, 1, False, 1, False, 2, 1)
        self.main = nn.Sequential(cat_module, block_block, tail)

    def forward(self, function_input):
        self._iterate_spectral_norms()
        return self.main(function_input)
