"""ctypes binding of libsg2_b200.so (the C ABI declared in include/sg2_b200.h).

The library is prebuilt in-tree by ``build.py`` (nvcc, sm_100a); there is no import-time JIT
(the reference blocks ~90 s in ``torch.utils.cpp_extension.load`` at import, op/fused_act.py:9-15).
There is NO fallback: if the library is missing, importing the ops raises, and calling an op on a
non-CUDA tensor raises ``RuntimeError('input must be a CUDA tensor')`` exactly like the reference's
``CHECK_CUDA`` (op/fused_bias_act.cpp:7,13-14).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# SG2_B200_LIB selects another build of the same ABI (A/B runs of kernel variants on one GPU box)
LIB_PATH = os.environ.get("SG2_B200_LIB") or os.path.join(_HERE, "csrc", "libsg2_b200.so")

SG2_F32, SG2_F16, SG2_BF16 = 0, 1, 2
_DTYPES = {torch.float32: SG2_F32, torch.float16: SG2_F16, torch.bfloat16: SG2_BF16}

c_i64, c_int, c_float, c_void_p = C.c_int64, C.c_int, C.c_float, C.c_void_p


class ConvParams(C.Structure):
    """mirror of sg2_conv_params (include/sg2_b200.h)"""
    _fields_ = [("weight", c_void_p), ("mod_weight", c_void_p), ("mod_bias", c_void_p),
                ("noise_weight", c_void_p), ("act_bias", c_void_p),
                ("cin", C.c_int32), ("cout", C.c_int32), ("ksize", C.c_int32),
                ("upsample", C.c_int32), ("latent_index", C.c_int32), ("resolution", C.c_int32)]


# name -> (restype, argtypes); must list every symbol the header declares (tests check this)
SIGNATURES = {
    "sg2_abi_version": (c_int, []),
    "sg2_last_error": (C.c_char_p, []),
    "sg2_launch_count": (c_i64, []),
    "sg2_note_launches": (None, [c_i64]),
    "sg2_selftest_fastdiv": (c_int, [C.c_uint32, C.c_uint32]),
    "sg2_fused_bias_act": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_i64, c_i64, c_i64, c_int,
                                   c_int, c_float, c_float, c_int, c_void_p]),
    "sg2_bias_act_grad_bias": (c_int, [c_void_p, c_void_p, c_i64, c_i64, c_i64, c_int, c_void_p]),
    "sg2_upfirdn2d": (c_int, [c_void_p, c_void_p, c_void_p, c_i64] + [c_int] * 14 + [c_void_p]),
    "sg2_equal_linear_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_i64, c_int, c_int,
                                     c_float, c_float, c_int, c_int, c_void_p]),
    "sg2_mapping_fwd": (c_int, [c_void_p, c_void_p, C.POINTER(c_void_p), C.POINTER(c_void_p), c_int,
                                c_i64, c_int, c_float, c_int, c_void_p, c_int, c_void_p]),
    "sg2_modconv2d_prep": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_float, c_int, c_void_p]),
    "sg2_modulation_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_i64, c_void_p, c_void_p, c_void_p,
                                   c_i64, c_int, c_int, c_int, c_float, c_float, c_int, c_void_p]),
    "sg2_modconv2d_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_i64, c_int, c_int,
                                  c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "sg2_noise_bias_act": (c_int, [c_void_p, c_void_p, c_void_p, c_i64, c_void_p, c_void_p, c_i64, c_int,
                                   c_i64, c_int, c_float, c_float, c_int, c_void_p]),
    "sg2_torgb_combine": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                                  c_int, c_i64, c_int, c_int, c_int, c_int, c_void_p]),
    "sg2_smooth_upsample2x": (c_int, [c_void_p, c_void_p, c_void_p, c_i64, c_int, c_int, c_int, c_void_p, c_i64,
                                      c_void_p, c_void_p, c_void_p, c_int, c_float, c_float, c_float, c_int, c_void_p]),
    "sg2_smooth_upsample2x_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_i64, c_int, c_int, c_int, c_void_p]),
    "sg2_ada_bias_act": (c_int, [c_void_p, c_void_p, c_void_p, c_i64, c_void_p, c_void_p, c_i64, c_int, c_i64,
                                 c_int, c_float, c_float, c_float, c_int, c_void_p]),
    "sg2_avg_pool_int": (c_int, [c_void_p, c_void_p, c_i64, c_int, c_int, c_int, c_int, c_void_p]),
    "sg2_resize_bilinear": (c_int, [c_void_p, c_void_p, c_i64, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "sg2_avg_pool_int_bwd": (c_int, [c_void_p, c_void_p, c_i64, c_int, c_int, c_int, c_int, c_void_p]),
    "sg2_resize_bilinear_bwd": (c_int, [c_void_p, c_void_p, c_i64, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "sg2_image_to_uint8": (c_int, [c_void_p, c_void_p, c_i64, c_int, c_void_p]),
    "sg2_synth_create": (c_int, [C.POINTER(c_void_p), c_int, c_int, c_int, C.POINTER(ConvParams), c_int,
                                 c_void_p, C.POINTER(c_float)]),
    "sg2_synth_create_ada": (c_int, [C.POINTER(c_void_p), c_int, c_int, c_int, C.POINTER(ConvParams), c_int,
                                     c_void_p, C.POINTER(c_float)]),
    "sg2_synth_enable_training": (c_int, [c_void_p]),
    "sg2_synth_backward": (c_int, [c_void_p, c_void_p, c_i64, C.POINTER(c_void_p), C.POINTER(c_i64), c_void_p, c_void_p, c_void_p]),
    "sg2_synth_destroy": (None, [c_void_p]),
    "sg2_synth_workspace_bytes": (c_i64, [c_void_p]),
    "sg2_synth_describe": (c_int, [c_void_p, C.c_char_p, c_int]),
    "sg2_synth_pack": (c_int, [c_void_p, c_void_p, c_void_p]),
    "sg2_conv3x3_tc_pack": (c_int, [c_void_p, c_void_p, c_int, c_int, C.c_float, c_void_p]),
    "sg2_conv3x3_tc": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_i64, c_int, c_int, c_int, c_void_p]),
    "sg2_nchw_to_nhwc_bf16": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_i64, c_int, c_i64, c_int, c_void_p]),
    "sg2_nhwc_bf16_to_nchw": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_i64, c_int, c_i64, c_int, c_void_p]),
    "sg2_rgb_modconv_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_i64, c_int, c_int, c_i64, c_int, c_void_p]),
    "sg2_rgb_modconv_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_i64, c_int, c_int, c_i64, c_int,
                                    c_void_p]),
    "sg2_nchw_to_polyphase_bf16": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_i64, c_int, c_int, c_int, c_void_p]),
    "sg2_polyphase_bf16_to_nchw": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_i64, c_int, c_int, c_int, c_void_p]),
    "sg2_sum_parts_bf16_to_nchw": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_i64, c_int, c_int, c_int,
                                           c_void_p]),
    "sg2_nhwc_bf16_to_nchw_act": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_i64, c_void_p, c_void_p, c_float, c_float,
                                          c_i64, c_int, c_i64, c_int, c_void_p]),
    "sg2_nchw_to_nhwc_bf16_actgrad": (c_int, [c_void_p, c_void_p, c_void_p, c_float, c_float, c_void_p, c_void_p, c_void_p,
                                              c_void_p, c_i64, c_int, c_i64, c_int, c_void_p]),
    "sg2_conv_taps_tc": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_i64, c_int, c_int, c_int, C.POINTER(c_int), c_int,
                                 c_void_p]),
    "sg2_conv_transpose3x3_tc": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_i64, c_int, c_int, c_int, c_void_p]),
    "sg2_synth_forward": (c_int, [c_void_p, c_void_p, c_void_p, c_i64, C.POINTER(c_void_p),
                                  C.POINTER(c_i64), c_void_p, c_void_p]),
    "sg2_synth_set_pooled_output": (c_int, [c_void_p, c_void_p, c_int, c_int]),
    "sg2_synth_set_profile_events": (c_int, [c_void_p, C.POINTER(c_void_p), c_int]),
    "sg2_synth_profile_events_used": (c_int, [c_void_p]),
}

_lib: Optional[C.CDLL] = None


def load() -> C.CDLL:
    """Load the prebuilt library (once) and attach prototypes.  Raises if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is not built. Run `python -m build_sg2` / `python __graft_entry__.py build` "
            "(nvcc, sm_100a). There is no CPU or PyTorch fallback for this path.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)   # AttributeError here = header / library mismatch
        fn.restype, fn.argtypes = res, args
    if lib.sg2_abi_version() != 1:
        raise ImportError(f"libsg2_b200 ABI version {lib.sg2_abi_version()} != 1")
    _lib = lib
    return lib


def check(status: int, what: str = "") -> None:
    if status != 0:
        msg = load().sg2_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"sg2_b200 {what}: {msg} (status {status})")


def dtype_code(t: torch.Tensor) -> int:
    try:
        return _DTYPES[t.dtype]
    except KeyError:
        raise RuntimeError(f"sg2_b200: unsupported dtype {t.dtype} (supported: float32, float16, bfloat16)") from None


def require_cuda(t: torch.Tensor, name: str = "input") -> None:
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor")


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def stream_of(t: torch.Tensor) -> int:
    """cudaStream_t of torch's current stream on the tensor's device (the reference uses the
    current device's stream without a device guard, fused_bias_act_kernel.cu:54-56)."""
    return torch.cuda.current_stream(t.device).cuda_stream


class device_of:
    """Explicit device guard around a C call (kernels launch on the tensor's device)."""

    def __init__(self, t: torch.Tensor):
        self.guard = torch.cuda.device(t.device)

    def __enter__(self):
        self.guard.__enter__()
        return self

    def __exit__(self, *exc):
        return self.guard.__exit__(*exc)


def launch_count() -> int:
    return int(load().sg2_launch_count())
