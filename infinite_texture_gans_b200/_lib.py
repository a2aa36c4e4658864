"""ctypes binding of libitg_b200.so (include/itg.h).

The library is the only compute backend of the package.  There is no CPU fallback: importing works
without a GPU (so the symbol table can be checked), but every compute entry point needs CUDA tensors,
and a missing library raises at first use with the build command.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ITG_B200_LIB") or os.path.join(_HERE, "libitg_b200.so")     # (override: A/B runs of two builds of the library)

# enums of include/itg.h
F32, F16, BF16 = 0, 1, 2
CONV3X3, CONV1X1, UPCONV = 0, 1, 2
BORDER_NONE, BORDER_REPLICATE, BORDER_CONSTANT = 0, 1, 2
RES_NONE, RES_GRID, RES_F32 = 0, 1, 2
IMPL_AUTO, IMPL_DIRECT, IMPL_UMMA, IMPL_TILE, IMPL_PAIR, IMPL_SPLIT = 0, 1, 2, 3, 4, 5
IMG_MERGED, IMG_PATCHES = 0, 1

DTYPE_OF = {torch.float32: F32, torch.float16: F16, torch.bfloat16: BF16}
TORCH_OF = {"f32": torch.float32, "f16": torch.float16, "bf16": torch.bfloat16}

EXPORTS = ("itg_version", "itg_last_error", "itg_conv_desc_size", "itg_conv_fwd", "itg_attention_fwd",
           "itg_pack_nchw", "itg_pack_map_taps", "itg_copy_rect", "itg_fill_frame", "itg_halo_exchange", "itg_step_advance",
           "itg_ipc_alloc", "itg_ipc_open", "itg_ipc_close", "itg_ipc_free", "itg_image_to_u8", "itg_ssm_fwd", "itg_ssm_desc_size", "itg_noise_normal")


class ConvDesc(C.Structure):
    """Mirror of `struct itg_conv_desc` (include/itg.h), same field order."""
    _fields_ = [
        ("dtype", C.c_int32), ("mode", C.c_int32), ("impl", C.c_int32), ("border", C.c_int32),
        ("in_", C.c_void_p), ("in_h", C.c_int32), ("in_w", C.c_int32), ("in_pitch", C.c_int32),
        ("in_c", C.c_int32), ("in_c_off", C.c_int32), ("k", C.c_int32),
        ("w", C.c_void_p), ("n_pad", C.c_int32), ("k_pad", C.c_int32), ("bias", C.c_void_p),
        ("out_h", C.c_int32), ("out_w", C.c_int32), ("out_c", C.c_int32),
        ("res_kind", C.c_int32), ("res_shift", C.c_int32), ("res", C.c_void_p),
        ("res_c", C.c_int32), ("res_h", C.c_int32), ("res_w", C.c_int32),
        ("mod_x", C.c_void_p), ("mod_c", C.c_int32), ("mod_shift", C.c_int32), ("mod_h", C.c_int32),
        ("mod_w", C.c_int32), ("mod_mean", C.c_void_p), ("mod_rstd", C.c_void_p),
        ("out_raw", C.c_void_p), ("out_act", C.c_void_p), ("scale", C.c_void_p), ("shift", C.c_void_p),
        ("leak", C.c_float), ("act_linear", C.c_int32), ("out_f32", C.c_void_p), ("out_img", C.c_void_p),
        ("img_c", C.c_int32), ("img_layout", C.c_int32), ("patch", C.c_int32),
        ("in2", C.c_void_p), ("in2_c", C.c_int32), ("in2_c_off", C.c_int32), ("k2", C.c_int32), ("w2", C.c_void_p), ("k2_pad", C.c_int32),
    ]


class SsmDesc(C.Structure):
    """Mirror of `struct itg_ssm_desc` (include/itg.h), same field order."""
    _fields_ = [
        ("dtype", C.c_int32), ("border", C.c_int32), ("h", C.c_int32), ("w", C.c_int32), ("c", C.c_int32), ("n_pad", C.c_int32),
        ("map", C.c_void_p), ("map_pitch", C.c_int32), ("x_shift", C.c_int32),
        ("w_mlp", C.c_void_p), ("w_embed", C.c_void_p), ("b_embed", C.c_void_p),
        ("x", C.c_void_p), ("x_c", C.c_int32), ("x_h", C.c_int32), ("x_w", C.c_int32), ("linear", C.c_int32),
        ("mean", C.c_void_p), ("rstd", C.c_void_p), ("out", C.c_void_p), ("leak", C.c_float), ("zero_ring", C.c_int32),
    ]


class ItgError(RuntimeError):
    pass


_lib: Optional[C.CDLL] = None


def load() -> C.CDLL:
    """Load libitg_b200.so (built in-tree by __graft_entry__.build()); raise loudly if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ItgError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc -gencode arch=compute_100a,code=sm_100a).  There is no CPU or PyTorch fallback for this path.")
    lib = C.CDLL(LIB_PATH)
    lib.itg_version.restype = C.c_int
    lib.itg_last_error.restype = C.c_char_p
    lib.itg_conv_desc_size.restype = C.c_int
    lib.itg_conv_fwd.restype = C.c_int
    lib.itg_conv_fwd.argtypes = [C.POINTER(ConvDesc), C.c_void_p]
    lib.itg_attention_fwd.restype = C.c_int
    lib.itg_attention_fwd.argtypes = ([C.c_int32, C.c_void_p] + [C.c_int32] * 5 + [C.c_void_p] * 9 +
                                      [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_int32, C.c_void_p])
    lib.itg_pack_nchw.restype = C.c_int
    lib.itg_pack_nchw.argtypes = [C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p]
    lib.itg_pack_map_taps.restype = C.c_int
    lib.itg_pack_map_taps.argtypes = [C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p]
    lib.itg_copy_rect.restype = C.c_int
    lib.itg_copy_rect.argtypes = [C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p] + [C.c_int32] * 6 + [C.c_void_p]
    lib.itg_halo_exchange.restype = C.c_int
    lib.itg_halo_exchange.argtypes = [C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_int32] + [C.c_void_p] * 9 + [C.c_int32, C.c_void_p]
    lib.itg_step_advance.restype = C.c_int
    lib.itg_step_advance.argtypes = [C.c_void_p, C.c_void_p]
    lib.itg_ipc_alloc.restype = C.c_int
    lib.itg_ipc_alloc.argtypes = [C.c_int32, C.c_uint64, C.POINTER(C.c_void_p), C.c_void_p]
    lib.itg_ipc_open.restype = C.c_int
    lib.itg_ipc_open.argtypes = [C.c_int32, C.c_void_p, C.POINTER(C.c_void_p)]
    lib.itg_ipc_close.restype = C.c_int
    lib.itg_ipc_close.argtypes = [C.c_void_p]
    lib.itg_ipc_free.restype = C.c_int
    lib.itg_ipc_free.argtypes = [C.c_void_p]
    lib.itg_fill_frame.restype = C.c_int
    lib.itg_fill_frame.argtypes = [C.c_int32, C.c_void_p] + [C.c_int32] * 5 + [C.c_void_p]
    lib.itg_image_to_u8.restype = C.c_int
    lib.itg_image_to_u8.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p]
    lib.itg_noise_normal.restype = C.c_int
    lib.itg_noise_normal.argtypes = [C.c_void_p] + [C.c_int32] * 7 + [C.c_uint64, C.c_uint32, C.c_void_p]
    lib.itg_ssm_desc_size.restype = C.c_int
    lib.itg_ssm_fwd.restype = C.c_int
    lib.itg_ssm_fwd.argtypes = [C.POINTER(SsmDesc), C.c_void_p]
    if lib.itg_ssm_desc_size() != C.sizeof(SsmDesc):
        raise ItgError(f"itg_ssm_desc layout mismatch: library {lib.itg_ssm_desc_size()} B, binding {C.sizeof(SsmDesc)} B")
    if lib.itg_conv_desc_size() != C.sizeof(ConvDesc):
        raise ItgError(f"itg_conv_desc layout mismatch: library {lib.itg_conv_desc_size()} B, binding {C.sizeof(ConvDesc)} B")
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != 0:
        raise ItgError(f"libitg_b200 error {rc}: {load().itg_last_error().decode()}")


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    """Device pointer of a CUDA tensor (None -> NULL).  Refuses host tensors: no CPU path exists."""
    if t is None:
        return None
    if not t.is_cuda:
        raise ItgError("libitg_b200 only takes CUDA tensors (there is no CPU fallback)")
    return t.data_ptr()


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream
