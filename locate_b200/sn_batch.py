"""Batched spectral-norm power iteration for a whole model (lb_sn_power_iter_batched).

The reference runs one power iteration inside every SpectralNorm.forward (spectral_norm.py:57-59).  u, v and sigma
of a layer depend only on (W_bar, u), and every wrapper is called exactly once per model forward, so iterating all
layers up front gives identical results with 4 kernel launches per forward instead of 4 per layer.  Generator /
Discriminator call `run()` at the top of forward(); each SpectralNorm then finds its sigma in `_pre_sigma`.
"""
import ctypes

import numpy as np
import torch

from ._lib import call, ptr

_ROWS_PER_ITEM1 = 256        # rows x 256 columns per pass-1 work item (64 K elements)


class _LayerRec(ctypes.Structure):
    _fields_ = [("w", ctypes.c_uint64), ("u", ctypes.c_uint64), ("v", ctypes.c_uint64),
                ("height", ctypes.c_int32), ("width", ctypes.c_int32), ("t_off", ctypes.c_int32), ("s_off", ctypes.c_int32)]


class SpectralBatch:
    def __init__(self, model):
        from .layers import SpectralNorm
        self.modules = [m for m in model.modules() if isinstance(m, SpectralNorm)]
        self._key = None
        self._tables = None

    def _build(self):
        mods = self.modules
        recs = (_LayerRec * len(mods))()
        items1, items3 = [], []
        off = 0
        for l, m in enumerate(mods):
            w = m.module.weight_bar
            height, width = m.spec.sn_shape
            recs[l].w, recs[l].u, recs[l].v = w.data_ptr(), m.module.weight_u.data_ptr(), m.module.weight_v.data_ptr()
            recs[l].height, recs[l].width = height, width
            recs[l].t_off, recs[l].s_off = off, off + width
            off += width + height
            for c0 in range(0, width, 256):
                for r0 in range(0, height, _ROWS_PER_ITEM1):
                    items1.append((l, c0, r0, min(_ROWS_PER_ITEM1, height - r0)))
            for r0 in range(0, height, 8):
                items3.append((l, r0))
        dev = mods[0].module.weight_bar.device
        rec_bytes = np.frombuffer(bytes(recs), dtype=np.uint8).copy()
        self._tables = dict(
            layers=torch.from_numpy(rec_bytes).to(dev),
            items1=torch.tensor(items1, dtype=torch.int32, device=dev),
            items3=torch.tensor(items3, dtype=torch.int32, device=dev),
            n1=len(items1), n3=len(items3), scratch_floats=off,
            scratch=torch.empty(off, dtype=torch.float32, device=dev))

    def run(self):
        """Iterate every layer once; hand each SpectralNorm its [sigma, 1/sigma] slice."""
        mods = self.modules
        if not mods:
            return
        first = mods[0].module
        key = (first.weight_bar.data_ptr(), first.weight_u.data_ptr(), mods[-1].module.weight_bar.data_ptr(),
               first.weight_bar.device)
        if key != self._key:          # parameters were re-homed (.to(device), optimizer arena): rebuild the tables
            self._build()
            self._key = key
        t = self._tables
        sigma = torch.empty((len(mods), 2), dtype=torch.float32, device=first.weight_bar.device)   # fresh: saved for backward
        call("lb_sn_power_iter_batched", ptr(t["layers"]), len(mods), ptr(t["items1"]), t["n1"], ptr(t["items3"]), t["n3"],
             ptr(t["scratch"]), t["scratch_floats"], ptr(sigma))
        for l, m in enumerate(mods):
            for _ in range(m.power_iterations - 1):
                raise NotImplementedError("power_iterations > 1 with the batched iteration")
            m._pre_sigma = sigma[l]
