"""ctypes binding of the C-ABI library (include/locate_b200.h).

There is NO fallback: if liblocate_b200.so is missing or a call returns a non-zero status this module
raises.  Pointers are raw `tensor.data_ptr()` values; the stream is torch's current CUDA stream so the
kernels order with torch's allocator and with NCCL collectives issued through torch.distributed.
"""
import ctypes
import os
from ctypes import POINTER, Structure, c_double, c_float, c_int, c_int64, c_size_t, c_void_p

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "liblocate_b200.so")


class ConvGeom(Structure):
    """Mirror of lb_conv_geom."""
    _fields_ = [(n, c_int) for n in ("batch", "in_h", "in_w", "in_c", "out_h", "out_w", "out_c",
                                     "kh", "kw", "stride", "pad", "mode", "ld_in", "ld_out")] + \
               [(n, c_int64) for n in ("w_sk", "w_sn", "w_sty", "w_stx")]


P = c_void_p
_SIGNATURES = {
    "lb_version": ([], c_int),
    "lb_sm_arch": ([], c_int),
    "lb_last_launch_count": ([], c_int),
    "lb_reset_launch_count": ([], None),
    "lb_set_pdl": ([c_int], c_int),
    "lb_roottanh_fwd": ([P, P, c_size_t, c_int, c_int, P], c_int),
    "lb_roottanh_bwd": ([P, P, P, c_size_t, c_int, c_int, P], c_int),
    "lb_tanh_fwd": ([P, P, c_size_t, c_int, P], c_int),
    "lb_tanh_bwd": ([P, P, P, c_size_t, c_int, P], c_int),
    "lb_add": ([P, P, P, c_size_t, c_int, P], c_int),
    "lb_mul": ([P, P, P, c_size_t, c_int, P], c_int),
    "lb_hinge_fwd": ([P, P, c_size_t, P], c_int),
    "lb_hinge_bwd": ([P, P, P, c_size_t, P], c_int),
    "lb_stat_work_doubles": ([], c_size_t),
    "lb_norm_stats": ([P, c_size_t, P, P, c_int, P], c_int),
    "lb_norm_finalize": ([P, c_double, P, P], c_int),
    "lb_norm_apply": ([P, P, P, c_int, P, P, c_int, c_int, c_int, c_int, P], c_int),
    "lb_norm_apply_ex": ([P, P, P, c_int, P, P, P, P, c_int, c_int, c_int, c_int, P], c_int),
    "lb_norm_bwd_reduce": ([P, P, P, P, P, c_int, c_int, c_int, c_int, P], c_int),
    "lb_norm_bwd_finalize": ([P, P, P, c_int, P, c_int, c_int, P, P, P, P], c_int),
    "lb_norm_bwd_apply": ([P, P, P, P, c_int, P, P, P, c_int, c_int, c_int, c_int, P], c_int),
    "lb_gate_fwd": ([P, P, P, P, c_int, c_int, c_int, c_int, c_int, P], c_int),
    "lb_gate_fwd_stats": ([P, P, P, P, P, P, c_int, c_int, c_int, c_int, c_int, P], c_int),
    "lb_gate_bwd": ([P, P, P, P, P, P, P, P, c_int, c_int, c_int, c_int, c_int, c_int, P], c_int),
    "lb_sn_power_iter_work_floats": ([c_int, c_int], c_size_t),
    "lb_sn_power_iter": ([P, c_int, c_int, P, P, P, P, P], c_int),
    "lb_sn_power_iter_batched": ([P, c_int, P, c_int, P, c_int, P, P, P, P], c_int),
    "lb_sn_weight_grad": ([P, P, P, P, P, P, c_int, c_int, c_int, P, P, P, P, P, P], c_int),
    "lb_sn_uv_grad_batched": ([P, c_int, P, c_int, P, P, P], c_int),
    "lb_wgrad_tc_supported": ([POINTER(ConvGeom)], c_int),
    "lb_augment": ([P, P, P, P, c_int, c_int, c_int, c_int, P], c_int),
    "lb_wgrad_tc_workspace_floats": ([POINTER(ConvGeom)], c_size_t),
    "lb_wgrad_tc": ([P, P, P, POINTER(ConvGeom), P, c_size_t, P, P, P, P], c_int),
    "lb_conv_gemm": ([P, P, P, P, P, POINTER(ConvGeom), P], c_int),
    "lb_conv_wgrad": ([P, P, P, POINTER(ConvGeom), P], c_int),
    "lb_conv_small_supported": ([POINTER(ConvGeom)], c_int),
    "lb_conv_small": ([P, P, P, P, P, POINTER(ConvGeom), c_int, P, c_int, c_int, c_int, c_int, P], c_int),
    "lb_conv_small_wgrad_supported": ([POINTER(ConvGeom)], c_int),
    "lb_conv_small_wgrad": ([P, P, P, POINTER(ConvGeom), c_int, c_int, P], c_int),
    "lb_dw_conv": ([P, P, P, P] + [c_int] * 12 + [P], c_int),
    "lb_dw_wgrad": ([P, P, P] + [c_int] * 11 + [P], c_int),
    "lb_gfull_fwd": ([P, P, P, P, c_int, c_int, c_int, c_int, c_int, P], c_int),
    "lb_gfull_dgrad": ([P, P, P, P, c_int, c_int, c_int, c_int, c_int, P], c_int),
    "lb_gfull_wgrad": ([P, P, P, c_int, c_int, c_int, c_int, c_int, P], c_int),
    "lb_conv_tc_supported": ([POINTER(ConvGeom)], c_int),
    "lb_conv_tc_packed_elems": ([POINTER(ConvGeom)], c_size_t),
    "lb_conv_tc_pack": ([P, P, POINTER(ConvGeom), P], c_int),
    "lb_pack_rec_bytes": ([], c_int),
    "lb_pack_chunk_items": ([], c_int),
    "lb_pack_rec_fill": ([P, P, POINTER(ConvGeom), P], c_int),
    "lb_conv_tc_pack_batched": ([P, P, c_int, P], c_int),
    "lb_conv_tc_gemm": ([P, P, P, P, P, POINTER(ConvGeom), P], c_int),
    "lb_conv_tc_workspace_bytes": ([POINTER(ConvGeom)], c_size_t),
    "lb_conv_tc_gemm_ws": ([P, P, P, P, P, POINTER(ConvGeom), P, c_size_t, c_int, P], c_int),
    "lb_conv_tc_ex_supported": ([POINTER(ConvGeom), c_int, c_int, c_int, c_int], c_int),
    "lb_conv_tc_gemm_ex": ([P, P, P, P, P, P, P, c_int, P, c_int, c_int, c_int, POINTER(ConvGeom), P], c_int),
    "lb_cast_bf16": ([P, P, c_size_t, P], c_int),
    "lb_cast_bf16_rows": ([P, c_int, P, c_int, c_int64, c_int, c_int, P], c_int),
    "lb_roottanh_fwd_bf16": ([P, P, c_size_t, c_int, P], c_int),
    "lb_colsum": ([P, c_int64, c_int, c_int, P, c_int, P], c_int),
    "lb_softmax_pixels_work_floats": ([c_int, c_int, c_int], c_size_t),
    "lb_softmax_pixels_fwd": ([P, P, c_int, c_int, c_int, P, c_size_t, c_int, P], c_int),
    "lb_softmax_pixels_bwd": ([P, P, P, c_int, c_int, c_int, P, c_size_t, c_int, P], c_int),
    "lb_softmax_rows_fwd": ([P, P, c_int, c_int, c_int, P], c_int),
    "lb_softmax_rows_bwd": ([P, P, P, c_int, c_int, c_int, P], c_int),
    "lb_featpool_fwd": ([P, P, c_int, c_int, c_int, c_int, c_int, c_int, P], c_int),
    "lb_featpool_bwd": ([P, P, c_int, c_int, c_int, c_int, c_int, c_int, P], c_int),
    "lb_upsample2x_fwd": ([P, P, c_int, c_int, c_int, c_int, c_int, P], c_int),
    "lb_upsample2x_bwd": ([P, P, c_int, c_int, c_int, c_int, c_int, P], c_int),
    "lb_avgpool2_fwd": ([P, P, c_int, c_int, c_int, c_int, c_int, P], c_int),
    "lb_avgpool2_bwd": ([P, P, c_int, c_int, c_int, c_int, c_int, P], c_int),
    "lb_copy_rows": ([P, c_int, P, c_int, c_int64, c_int, c_int, c_int, c_int, P], c_int),
    "lb_nchw_to_nhwc": ([P, P, c_int, c_int, c_int, c_int, P], c_int),
    "lb_nhwc_to_nchw": ([P, P, c_int, c_int, c_int, c_int, P], c_int),
    "lb_loss_sums": ([P, P, c_int, P, P], c_int),
    "lb_d_loss": ([P, P, P, P, c_int, c_double, c_float, P, P, P, P, P], c_int),
    "lb_g_loss": ([P, c_int, c_double, P, P, P], c_int),
    "lb_nadam_step": ([P, P, P, P, c_size_t, c_float, c_float, c_float, P, P], c_int),
    "lb_nadam_schedule": ([P, P, c_double, c_double, c_double, c_double, P], c_int),
    "lb_fill": ([P, c_size_t, c_float, P], c_int),
    "lb_scale": ([P, c_size_t, c_float, P], c_int),
}

_ERRORS = {-1: "LB_EINVAL (bad size / null pointer)", -2: "LB_EALIGN", -3: "LB_EUNSUPPORTED"}


class LocateLibraryError(RuntimeError):
    pass


def exported_symbols():
    return sorted(_SIGNATURES)


def _load():
    if not os.path.exists(LIB_PATH):
        raise LocateLibraryError(
            f"{LIB_PATH} is missing: build it with `python -m locate_b200.build` "
            "(locate_b200 has no CPU or PyTorch fallback path)")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (argtypes, restype) in _SIGNATURES.items():
        fn = getattr(lib, name)           # AttributeError if the library does not export it
        fn.argtypes = argtypes
        fn.restype = restype
    return lib


_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        _LIB = _load()
    return _LIB


def stream_ptr():
    return torch.cuda.current_stream().cuda_stream


def ptr(t):
    """Device pointer of a tensor (None -> NULL).  The product path is CUDA-only."""
    if t is None:
        return None
    if not t.is_cuda:
        raise LocateLibraryError("locate_b200 kernels need CUDA tensors (no CPU fallback)")
    return t.data_ptr()


def call(name, *args):
    """Invoke an entry point on torch's current stream; raise on any non-zero status."""
    rc = getattr(lib(), name)(*args, stream_ptr())
    if rc != 0:
        detail = _ERRORS.get(rc, f"cudaError {rc}" if rc > 0 else f"status {rc}")
        raise LocateLibraryError(f"{name} failed: {detail}")


def launch_count():
    return lib().lb_last_launch_count()


def reset_launch_count():
    lib().lb_reset_launch_count()


def set_pdl(on):
    """Programmatic dependent launch of the library's kernels on / off (include/locate_b200.h lb_set_pdl); returns
    the previous setting.  Takes effect for launches made afterwards (a captured graph keeps the edges it was
    captured with)."""
    return bool(lib().lb_set_pdl(1 if on else 0))
