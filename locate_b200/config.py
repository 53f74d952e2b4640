"""Configuration with the reference's constant names (libs/config.py:19-73) as a settable object.

The reference freezes sizes at import time; here `configure(IMAGE_SIZE=32)` changes the values the
constructors (Generator(), Discriminator(), Block...) read when they are CALLED.
"""
import dataclasses
import math
import os


@dataclasses.dataclass
class Config:
    IMAGE_SIZE: int = 128
    BATCH_SIZE: int = 16
    MINIBATCHES: int = 8
    DITERS: int = 1
    END_LAYER: int = 1
    START_LAYER: int = 0
    FACTOR: int = 2
    G_STRIDE: int = 2
    D_STRIDE: int = 2
    D_HINGE: bool = True
    G_HINGE: bool = True
    SEPARABLE: bool = False
    FEATURE_MULTIPLIER: int = 1
    ROOTTANH_GROWTH: int = 4
    BASE_FEATURE_FACTOR: int = 8
    BOTTLENECK: int = 4
    MIN_ATTENTION_SIZE: int = 8
    ATTENTION_EVERY_NTH_LAYER: int = 2
    DEPTH: int = 1
    GLR: float = 5e-4
    DLR: float = 2e-3
    BETA_1: float = 0.5
    BETA_2: float = 0.9
    STRICT_REFERENCE: bool = True      # reproduce the reference's d-gamma = sum(x*x*g) (merge.py:33-38)
    # "bf16": conv/linear GEMMs on the tcgen05 tensor cores (bf16 operands, fp32 accumulate) wherever the
    # geometry allows; "fp32": every GEMM on the fp32 SIMT kernel (bit-for-bit the reference's precision class)
    PRECISION: str = "bf16"
    SMALL_KERNELS: bool = os.environ.get("LB_SMALL_KERNELS", "1") != "0"         # direct fp32 kernels for layers with <= 4 channels on one side (lb_conv_small*)

    @property
    def LAYERS(self):
        return int(math.log(self.IMAGE_SIZE, 2))

    @property
    def INPUT_VECTOR_Z(self):
        return self.IMAGE_SIZE

    @property
    def GEN_FEATURES(self):
        return self.FACTOR ** int(math.log(self.IMAGE_SIZE, self.G_STRIDE)) * self.BASE_FEATURE_FACTOR * 3

    @property
    def DIS_FEATURES(self):
        return self.FACTOR ** int(math.log(self.IMAGE_SIZE, self.D_STRIDE)) * self.BASE_FEATURE_FACTOR


CFG = Config()


def configure(**kwargs):
    for k, v in kwargs.items():
        if not hasattr(CFG, k) or isinstance(getattr(type(CFG), k, None), property):
            raise AttributeError(f"unknown config constant {k}")
        setattr(CFG, k, v)
    return CFG


def reset():
    configure(**{f.name: f.default for f in dataclasses.fields(Config)})
