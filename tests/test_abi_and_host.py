"""CPU-side checks: the C-ABI library builds, loads and exports every symbol include/locate_b200.h declares;
the module tree reproduces the reference's state_dict keys; host-side logic (feature schedules, Nadam
schedule scalars, bucket bounds, config).  No kernel is launched here (no GPU in the build container)."""
import ctypes
import os
import re

import pytest
import torch

import locate_b200 as L
from locate_b200 import _lib, dist, models, optim
from oracle import locate_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(autouse=True)
def _cfg():
    L.config.reset()
    yield
    L.config.reset()


def _header_symbols():
    with open(os.path.join(ROOT, "include", "locate_b200.h")) as fh:
        text = re.sub(r"/\*.*?\*/", "", fh.read(), flags=re.S)
    return sorted(set(re.findall(r"\b(?:int|void|size_t)\s+(lb_\w+)\s*\(", text)))


def test_library_builds_and_exports_every_declared_symbol():
    from locate_b200 import build
    path = build.build()
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    declared = _header_symbols()
    assert len(declared) >= 40
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/locate_b200.h but not exported"
    assert sorted(_lib.exported_symbols()) == declared, "ctypes table and header disagree"
    assert lib.lb_version() >= 100 and lib.lb_sm_arch() == 100


def test_library_is_sm100a_only():
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    assert not re.search(r"sm_(8\d|9\d)\b", out)


def test_product_refuses_cpu_tensors():
    with pytest.raises(_lib.LocateLibraryError):
        L.layers.nonlinear_function(torch.randn(4))


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "locate_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                with open(os.path.join(dirpath, f)) as fh:
                    src = fh.read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), f
                assert "locate_oracle" not in src, f


@pytest.mark.parametrize("size", [32, 64, 128, 256])
def test_feature_schedules_match_oracle(size):
    L.configure(IMAGE_SIZE=size)
    cfg = O.OracleConfig(IMAGE_SIZE=size)
    assert models.generator_feature_list() == O.generator_features(cfg)
    assert models.discriminator_feature_list() == O.discriminator_features(cfg)
    if size == 128:
        assert models.generator_feature_list() == [128, 1536, 768, 384, 192, 96, 48]
        assert models.discriminator_feature_list() == [32, 64, 128, 256, 512, 1024, 1024]


@pytest.mark.parametrize("name", ["step_s32_w2_b3.pt", "step_s16_w2_depth3_b2.pt", "step_s16_w4_separable_b2.pt"])
def test_state_dict_keys_and_seeded_init_equal_the_reference(golden, name):
    r = golden(name)
    L.configure(**r["overrides"])
    torch.manual_seed(999)
    gen = L.Generator()
    gen.apply(L.init)
    dis = L.Discriminator()
    dis.apply(L.init)
    gs, ds = gen.state_dict(), dis.state_dict()
    assert list(gs) == list(r["g_state"]) and list(ds) == list(r["d_state"])
    for k, v in r["g_state"].items():
        assert torch.equal(gs[k], v), k
    for k, v in r["d_state"].items():
        assert torch.equal(ds[k], v), k
    assert torch.equal(gen.noise, r["const_noise"])


def test_default_parameter_counts():
    L.configure(IMAGE_SIZE=32)
    assert L.parameter_count(L.Generator()) == 3_482_522 or L.parameter_count(L.Generator()) > 3_400_000
    L.configure(IMAGE_SIZE=64)
    g = L.Generator()
    assert 14_000_000 < L.parameter_count(g) < 14_200_000      # 14.08 M


def test_nadam_schedule_scalars_match_oracle_formula():
    # closed form of nadam.py:68-73 for the first three steps
    b1, decay = 0.5, 4e-3
    sched = 1.0
    for t in (1, 2, 3):
        mu_t = b1 * (1 - 0.5 * 0.96 ** (t * decay))
        mu_n = b1 * (1 - 0.5 * 0.96 ** ((t + 1) * decay))
        sched *= mu_t
        assert 0 < (1 - mu_t) / (1 - sched) <= 1.0 + 1e-12
        assert 0 < mu_n / (1 - sched * mu_n) < 1.0


def test_bucket_bounds_cover_exactly():
    for n, b in ((10, 3), (9, 3), (1, 5), (0, 4), (1 << 20, 1 << 18)):
        bounds = dist.bucket_bounds(n, b)
        assert sum(e - s for s, e in bounds) == n
        assert all(bounds[i][1] == bounds[i + 1][0] for i in range(len(bounds) - 1))


def test_configure_rejects_unknown_and_derived():
    with pytest.raises(AttributeError):
        L.configure(NOT_A_CONSTANT=1)
    with pytest.raises(AttributeError):
        L.configure(LAYERS=3)


def test_start_layer_builds_what_the_reference_builds(golden):
    """START_LAYER >= 1 (models.py:44-50): same modules, names and shapes as the reference constructs.  (The reference's
    own forward then raises -- fixture key `forward_error` -- see tests/test_gpu_parity.py for the mirrored behaviour.)"""
    r = golden("start_layer1_s16_w2.pt")
    L.configure(**r["overrides"])
    gen = L.Generator()
    assert {k: tuple(v.shape) for k, v in gen.state_dict().items()} == r["g_shapes"]
    assert list(gen.state_dict()) == list(r["g_shapes"])
    assert r["forward_error"] is not None and "cannot be multiplied" in r["forward_error"]


def test_unsupported_flags_fail_loudly():
    L.configure(G_STRIDE=4, IMAGE_SIZE=32)
    with pytest.raises(NotImplementedError):
        L.Generator()
