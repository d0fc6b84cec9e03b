"""ctypes binding of the C-ABI library ``_C/libscenesplat_b200.so`` (include/scenesplat_b200.h).

There is deliberately NO fallback: if the library cannot be loaded, or a call is made without a CUDA
device, the product path raises.  torch is used only for device memory and streams.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_C", "libscenesplat_b200.so")

_vp, _i, _i64, _sz, _f, _d = C.c_void_p, C.c_int, C.c_int64, C.c_size_t, C.c_float, C.c_double

# name -> (restype, argtypes); mirrors include/scenesplat_b200.h one to one
SIGNATURES = {
    "ss_coord_max": (_i, [_vp, _i, _i64, _vp, _vp]),
    "ss_serialize_workspace_bytes": (_sz, [_i64, _i, _i, _i]),
    "ss_serialize": (_i, [_vp, _i, _vp, _i, _i64, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "ss_gridsample_workspace_bytes": (_sz, [_i64]),
    "ss_gridsample_index": (_i, [_vp, _i64, _d, _i, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "ss_gridsample_select": (_i, [_vp, _i64, _d, _vp, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp]),
    "ss_gather_rows": (_i, [_vp, _i64, _vp, _vp, _i64, _vp, _vp]),
    "ss_pool_workspace_bytes": (_sz, [_i64]),
    "ss_pool_index": (_i, [_vp, _vp, _vp, _vp, _i64, _i, _i, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                           _vp, _sz, _vp]),
    "ss_segment_reduce": (_i, [_vp, _i, _vp, _vp, _vp, _i64, _i, _i, _vp, _vp, _i, _vp, _i, _vp]),
    "ss_pool_reduce": (_i, [_vp, _i, _vp, _vp, _vp, _i64, _i, _i, _vp, _vp, _i, _vp, _i, _vp, _vp]),
    "ss_unpool_gather_add": (_i, [_vp, _vp, _i, _vp, _i64, _i, _vp, _vp, _vp, _vp, _i, _vp, _vp, _i, _vp]),
    "ss_kmap_workspace_bytes": (_sz, [_i64, _i]),
    "ss_kmap_build": (_i, [_vp, _i, _vp, _vp, _vp, _i64, _i, _i, _i, _vp, _vp, _vp, _sz, _vp]),
    "ss_kmap_subset": (_i, [_vp, _vp, _i64, _i, _i, _vp, _vp, _vp]),
    "ss_kmap_pairs": (_i, [_vp, _vp, _i64, _i, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "ss_subm_conv_simt": (_i, [_vp, _i, _vp, _vp, _vp, _vp, _vp, _i, _i64, _i, _i, _i, _vp, _i, _vp]),
    "ss_subm_conv_gemm256": (_i, [_vp, _vp, _vp, _vp, _i64, _i, _i, _i, _vp, _vp]),
    "ss_subm_conv_gemm_pair": (_i, [_vp, _vp, _vp, _vp, _i64, _i, _i, _i, _vp, _vp]),
    "ss_subm_conv_reduce": (_i, [_vp, _vp, _vp, _i64, _i, _i, _vp, _i, _vp]),
    "ss_subm_conv_reduce_add_ln": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _f, _i64, _i, _i, _vp, _vp, _vp]),
    "ss_subm_conv_fused_add_ln": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                       _vp, _vp, _f, _i64, _vp, _vp, _vp]),
    "ss_patch_table": (_i, [_vp, _i, _i, _i, _vp, _vp, _vp]),
    "ss_patch_attention_simt": (_i, [_vp, _i, _vp, _vp, _i, _i, _i, _i, _f, _vp, _i, _vp]),
    "ss_patch_attention": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _f, _vp, _vp]),
    "ss_linear_act_bf16": (_i, [_vp, _vp, _vp, _i64, _i, _i, _i, _vp, _vp]),
    "ss_linear_residual_bf16": (_i, [_vp, _vp, _vp, _vp, _i64, _i, _i, _vp, _vp, _vp]),
    "ss_stem_conv_wgrad_workspace_bytes": (_sz, [_i, _i]),
    "ss_stem_conv_wgrad": (_i, [_vp, _vp, _vp, _i64, _i, _i, _i, _vp, _vp, _sz, _vp]),
    "ss_colsum_workspace_bytes": (_sz, [_i]),
    "ss_colsum_bf16": (_i, [_vp, _i64, _i, _vp, _vp, _sz, _vp]),
    "ss_patch_attention_lse": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _f, _vp, _vp, _i64, _vp]),
    "ss_patch_attention_backward_workspace_bytes": (_sz, [_i64, _i, _i]),
    "ss_patch_attention_backward": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _f, _i64, _vp, _vp, _sz, _vp]),
    "ss_add_layernorm": (_i, [_vp, _vp, _i, _vp, _vp, _vp, _vp, _f, _i64, _i, _vp, _vp, _i, _vp]),
    "ss_add_l2_normalize": (_i, [_vp, _vp, _i, _f, _i64, _i, _vp, _vp, _vp]),
    "ss_affine_act": (_i, [_vp, _i, _vp, _vp, _i, _i64, _i, _vp, _i, _vp]),
    "ss_l2_normalize": (_i, [_vp, _i, _i64, _i, _f, _vp, _i, _vp]),
    "ss_lang_head": (_i, [_vp, _i, _vp, _i64, _i, _i, _i, _f, _i, _vp, _vp, _vp, _vp, _vp]),
    "ss_lang_head_tc": (_i, [_vp, _vp, _i64, _i, _i, _f, _i, _vp, _vp, _vp, _vp, _vp]),
    "ss_cos_l2_loss": (_i, [_vp, _i, _vp, _i, _vp, _i64, _i, _vp, _vp]),
    "ss_class_half_sums": (_i, [_vp, _i, _vp, _vp, _vp, _i64, _i, _i, _vp, _vp, _vp]),
    "ss_segment_mean_bwd": (_i, [_vp, _i, _vp, _vp, _i64, _i, _i, _vp, _i, _vp]),
    "ss_unpool_gather_add_bwd": (_i, [_vp, _i, _vp, _vp, _i64, _i, _vp, _i, _vp]),
    "ss_cos_l2_loss_bwd": (_i, [_vp, _i, _vp, _i, _vp, _i64, _i, _vp, _vp, _f, _f, _vp, _vp]),
    "ss_class_half_sums_bwd": (_i, [_vp, _vp, _vp, _vp, _i64, _i, _i, _vp, _vp]),
    "ss_gelu_backward_bf16": (_i, [_vp, _vp, _i64, _vp, _vp]),
    "ss_subm_conv_wgrad": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp]),
    "ss_layernorm_backward": (_i, [_vp, _i, _vp, _i, _vp, _f, _i64, _i, _vp, _vp, _vp, _vp]),
    "ss_sphere_crop_workspace_bytes": (_sz, [_i64]),
    "ss_sphere_crop_order": (_i, [_vp, _i64, _vp, _vp, _vp, _vp, _sz, _vp]),
    "ss_vote_bin_count": (_i, [_vp, _i64, _vp, _f, _i, _i, _i, _vp, _vp, _vp]),
    "ss_vote_bin_fill": (_i, [_vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp]),
    "ss_knn_vote": (_i, [_vp, _vp, _vp, _f, _i, _i, _i, _vp, _i64, _i, _i, _i, _vp, _vp]),
    "ss_confusion_update": (_i, [_vp, _vp, _i64, _i, _i64, _vp, _vp, _vp]),
    "ss_version": (C.c_char_p, []),
    "ss_launch_count": (C.c_uint64, []),
}

_lib = None


def load():
    """Load (building first if the .so is absent and nvcc exists).  Raises on failure -- no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        from . import build as _build
        _build.build()
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing: fail loudly
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class CudaKernelError(RuntimeError):
    pass


def check(rc: int, what: str):
    if rc != 0:
        if rc == -1:
            raise CudaKernelError(f"{what}: bad arguments")
        raise CudaKernelError(f"{what}: CUDA error {rc}")


def ptr(t):
    """Device pointer of a tensor (or None -> NULL) as a plain int.  Tensors must be contiguous CUDA tensors."""
    if t is None:
        return None
    if not t.is_cuda:
        raise CudaKernelError("scenesplat_b200 kernels need CUDA tensors (there is no CPU fallback)")
    if not t.is_contiguous():
        raise CudaKernelError("scenesplat_b200 kernels need contiguous tensors")
    return t.data_ptr()


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def stream():
    """cudaStream_t of torch's current stream (raw handle; ~0.3 us instead of ~20 us for current_stream())."""
    if _raw_stream is not None:
        return _raw_stream(torch.cuda.current_device())
    return torch.cuda.current_stream().cuda_stream


def int_array(vals):
    return (C.c_int * len(vals))(*[int(v) for v in vals])


def float_array(vals):
    """Small HOST array of floats passed by pointer (e.g. a grid origin)."""
    return C.cast((C.c_float * len(vals))(*[float(v) for v in vals]), C.c_void_p)


# Optional per-call device timing (bench.py): name -> list of (start_event, end_event, meta).
# PROFILE_ONLY (a set of names) restricts the instrumentation to a few kernels so it can stay on inside a timed region.
PROFILE = None
PROFILE_ONLY = None


_fn_cache = {}


def call(name: str, *args, meta=None):
    fn = _fn_cache.get(name)
    if fn is None:
        fn = _fn_cache[name] = getattr(load(), name)
    if PROFILE is None or (PROFILE_ONLY is not None and name not in PROFILE_ONLY):
        rc = fn(*args)
        if rc != 0:
            check(rc, name)
        return
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    rc = fn(*args)
    e1.record()
    check(rc, name)
    PROFILE.setdefault(name, []).append((e0, e1, meta))


def launch_count() -> int:
    return int(load().ss_launch_count())


def workspace(nbytes: int, device):
    return torch.empty(int(nbytes) + 256, dtype=torch.uint8, device=device)
