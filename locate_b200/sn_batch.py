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
_COLS_PER_ITEM1 = 1024      # k_snb_wt_u: 256 threads x 4 consecutive columns (one 16-byte load per row)


class _LayerRec(ctypes.Structure):
    _fields_ = [("w", ctypes.c_uint64), ("u", ctypes.c_uint64), ("v", ctypes.c_uint64), ("dv", ctypes.c_uint64),
                ("height", ctypes.c_int32), ("width", ctypes.c_int32), ("t_off", ctypes.c_int32), ("s_off", ctypes.c_int32),
                ("nsplit", ctypes.c_int32), ("rows_per_split", ctypes.c_int32)]


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
        off = s_off = 0
        self._s_slices = []
        for l, m in enumerate(mods):
            w = m.module.weight_bar
            height, width = m.spec.sn_shape
            recs[l].w, recs[l].u, recs[l].v = w.data_ptr(), m.module.weight_u.data_ptr(), m.module.weight_v.data_ptr()
            dv = getattr(m.module.weight_v, "_lb_grad", None)       # arena view (optim.Nadam): gradient of a TRAINABLE v
            recs[l].dv = dv.data_ptr() if dv is not None else 0
            nsplit = (height + _ROWS_PER_ITEM1 - 1) // _ROWS_PER_ITEM1
            recs[l].height, recs[l].width = height, width
            recs[l].nsplit, recs[l].rows_per_split = nsplit, _ROWS_PER_ITEM1
            # W^T u is staged as one partial row per row split (summed in a fixed order by the normalise kernel:
            # the iteration is bit-reproducible, like the reference's torch.mv), then W v
            recs[l].t_off, recs[l].s_off = off, s_off
            self._s_slices.append((s_off, height))
            off += nsplit * width
            s_off += height
            for c0 in range(0, width, _COLS_PER_ITEM1):
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
            n1=len(items1), n3=len(items3), s_floats=s_off,
            scratch=torch.empty(off, dtype=torch.float32, device=dev),
            cacc=torch.zeros(len(mods), dtype=torch.float32, device=dev))     # per layer: sum over backward passes of dL/dsigma
        self.uv_pending = False

    def run(self):
        """Iterate every layer once; hand each SpectralNorm its [sigma, 1/sigma] slice."""
        mods = self.modules
        if not mods:
            return
        first = mods[0].module
        dv0 = getattr(first.weight_v, "_lb_grad", None)
        key = (first.weight_bar.data_ptr(), first.weight_u.data_ptr(), mods[-1].module.weight_bar.data_ptr(),
               first.weight_bar.device, dv0.data_ptr() if dv0 is not None else 0)
        if key != self._key:          # parameters were re-homed (.to(device), optimizer arena): rebuild the tables
            self._build()
            self._key = key
        t = self._tables
        dev = first.weight_bar.device
        sigma = torch.empty((len(mods), 2), dtype=torch.float32, device=dev)   # fresh: saved for backward
        s_out = torch.empty(t["s_floats"], dtype=torch.float32, device=dev)    # fresh: W v of THIS pass (backward of a trainable u)
        call("lb_sn_power_iter_batched", ptr(t["layers"]), len(mods), ptr(t["items1"]), t["n1"], ptr(t["items3"]), t["n3"],
             ptr(t["scratch"]), ptr(s_out), ptr(sigma))
        for l, m in enumerate(mods):
            for _ in range(m.power_iterations - 1):
                raise NotImplementedError("power_iterations > 1 with the batched iteration")
            sig = sigma[l]
            off, height = self._s_slices[l]
            sig._lb_uv = (s_out[off:off + height], t["cacc"][l:l + 1], self)
            m._pre_sigma = sig

    def finish_uv_grads(self):
        """dv += C * W^T u for the layers whose u / v are trainable (the reference's main.py:172 quirk, see
        include/locate_b200.h lb_sn_weight_grad); a no-op unless a backward pass accumulated into `cacc` since."""
        if not self.uv_pending or self._tables is None:
            return
        self.uv_pending = False
        t = self._tables
        call("lb_sn_uv_grad_batched", ptr(t["layers"]), len(self.modules), ptr(t["items1"]), t["n1"], ptr(t["scratch"]),
             ptr(t["cacc"]))
