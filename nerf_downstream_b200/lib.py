"""ctypes binding of the C-ABI library `libsparseconv_b200.so` (include/sparseconv_b200.h).

Every compute entry point takes raw device pointers and a CUDA stream.  There is NO CPU
fallback: calling a compute function without the library or without CUDA tensors raises.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_int64, c_void_p
from pathlib import Path

import torch

PKG_DIR = Path(__file__).resolve().parent
LIB_PATH = PKG_DIR / "libsparseconv_b200.so"

PREC_FP32 = 0
PREC_TF32 = 1
PREC_BF16 = 2
SRC_FLOAT, SRC_INT, SRC_STRIDE = 0, 1, 2
SLOT_BYTES = 16

# name -> (restype, argtypes); mirrors include/sparseconv_b200.h one to one
_P = c_void_p
SIGNATURES = {
    "spc_abi_version": (c_int, []),
    "spc_last_error": (c_char_p, []),
    "spc_launch_count": (c_int64, []),
    "spc_table_slots": (c_int64, [c_int64]),
    "spc_coords_insert_workspace": (c_int64, [c_int64]),
    "spc_coords_insert": (c_int, [_P, c_int64, c_int, _P, _P, c_int64, _P, _P, _P, _P, _P, _P, c_int64, _P]),
    "spc_coords_insert_dev": (c_int, [_P, c_int64, _P, c_int, _P, _P, c_int64, _P, _P, _P, _P, _P, _P, c_int64, _P]),
    "spc_kernel_map": (c_int, [_P, c_int64, _P, c_int64, _P, c_int, _P, _P, _P]),
    "spc_kernel_map_sym": (c_int, [_P, c_int64, _P, c_int64, _P, c_int, _P, _P, _P]),
    "spc_tile_mask": (c_int, [_P, c_int64, c_int, _P, _P]),
    "spc_kernel_map_transpose": (c_int, [_P, c_int64, c_int64, c_int, _P, _P]),
    "spc_pairs_workspace": (c_int64, [c_int64, c_int]),
    "spc_kernel_map_pairs": (c_int, [_P, c_int64, c_int, c_int64, _P, _P, _P, c_int64, _P]),
    "spc_segment_reduce": (c_int, [_P, _P, _P, c_int64, c_int64, c_int, c_int, _P, _P]),
    "spc_gather_rows": (c_int, [_P, _P, _P, c_int64, c_int, _P, _P]),
    "spc_scatter_add_rows": (c_int, [_P, _P, c_int64, c_int64, c_int, _P, _P]),
    "spc_row_masks": (c_int, [_P, c_int64, c_int, _P, _P, _P]),
    "spc_table_relabel": (c_int, [_P, c_int64, _P, _P]),
    "spc_debug_force_mt": (None, [c_int]),
    "spc_debug_set": (None, [c_int, c_int]),
    "spc_conv_path_counts": (None, [_P, c_int]),
    "spc_conv_tensor_core": (c_int, [c_int, c_int, c_int, c_int, c_int]),
    "spc_conv_packed_bytes": (c_int64, [c_int, c_int, c_int]),
    "spc_conv_pack_weights": (c_int, [_P, c_int, c_int, c_int, c_int, c_int, _P, _P]),
    "spc_conv_pack_weights_batch": (c_int, [_P, c_int, _P]),
    "spc_conv_fwd_packed": (c_int, [_P, _P, _P, _P, _P, c_int64, c_int64, c_int, c_int, c_int, c_int, _P, _P]),
    "spc_conv_dgrad_packed": (c_int, [_P, _P, _P, _P, c_int64, c_int64, c_int, c_int, c_int, c_int, _P, _P]),
    "spc_conv_dgrad_packed_acc": (c_int, [_P, _P, _P, _P, c_int64, c_int64, c_int, c_int, c_int, c_int, _P, c_int, _P]),
    "spc_conv_wgrad_acc": (c_int, [_P, _P, _P, _P, c_int64, c_int64, c_int, c_int, c_int, c_int, _P, c_int, _P]),
    "spc_conv_workspace": (c_int64, [c_int, c_int, c_int, c_int]),
    "spc_to_bf16": (c_int, [_P, c_int64, c_int, c_int64, c_int, _P, _P]),
    "spc_conv_fwd": (c_int, [_P, _P, _P, _P, _P, c_int64, c_int64, c_int, c_int, c_int, c_int, _P, _P, c_int64, _P]),
    "spc_conv_fwd_stats": (c_int, [_P, _P, _P, _P, _P, c_int64, c_int64, c_int, c_int, c_int, c_int, _P, _P, _P, _P, c_int64, _P]),
    "spc_bn_finalize": (c_int, [_P, c_int64, c_int, _P, _P, _P, _P, c_float, _P, _P]),
    "spc_conv_fwd_packed_stats": (c_int, [_P, _P, _P, _P, _P, c_int64, c_int64, c_int, c_int, c_int, c_int, _P, _P, _P, _P]),
    "spc_conv_dgrad": (c_int, [_P, _P, _P, _P, c_int64, c_int64, c_int, c_int, c_int, c_int, _P, _P, c_int64, _P]),
    "spc_conv_wgrad": (c_int, [_P, _P, _P, _P, c_int64, c_int64, c_int, c_int, c_int, c_int, _P, _P, c_int64, _P]),
    "spc_bn_workspace": (c_int64, [c_int64, c_int]),
    "spc_bn_stats": (c_int, [_P, c_int64, c_int, _P, _P, _P, _P, c_float, _P, c_int64, _P]),
    "spc_bn_apply": (c_int, [_P, _P, _P, _P, _P, _P, c_int64, c_int, c_float, c_int, _P, _P, _P]),
    "spc_bn_bwd": (c_int, [_P, _P, _P, _P, c_int64, _P, _P, _P, c_int64, c_int, c_float, c_int, c_int, _P, _P, _P, _P, _P, _P, c_int64, _P]),
    "spc_bn_bwd_acc": (c_int, [_P, _P, _P, _P, c_int64, _P, _P, _P, _P, c_int64, c_int, c_float, c_int, c_int, _P, _P, _P, _P, _P, c_int, _P, c_int64, _P]),
    "spc_bn_stats_tracked": (c_int, [_P, c_int64, c_int, _P, _P, _P, _P, c_float, _P, _P, c_int64, _P]),
    "spc_copy_rows": (c_int, [_P, c_int64, _P, c_int64, c_int64, c_int64, _P]),
    "spc_relu_fwd": (c_int, [_P, c_int64, _P, _P]),
    "spc_relu_bwd": (c_int, [_P, _P, c_int64, _P, _P]),
    "spc_add": (c_int, [_P, _P, c_int64, _P, _P]),
    "spc_pool_fwd": (c_int, [_P, _P, c_int64, c_int, c_int, c_int, _P, _P]),
    "spc_global_pool_fwd": (c_int, [_P, _P, c_int64, c_int, c_int, c_int, _P, _P, _P]),
    "spc_global_pool_bwd": (c_int, [_P, _P, _P, c_int64, c_int, c_int, c_int, _P, _P]),
    "spc_ce_fwd": (c_int, [_P, _P, c_int64, c_int, c_int64, _P, _P, _P, _P]),
    "spc_ce_bwd": (c_int, [_P, _P, _P, c_int64, c_int, _P, _P]),
    "spc_pool_max_fwd": (c_int, [_P, _P, c_int64, c_int, c_int, _P, _P, _P]),
    "spc_pool_max_bwd": (c_int, [_P, _P, c_int64, c_int64, c_int, _P, _P]),
    "spc_global_max_fwd": (c_int, [_P, _P, c_int64, c_int, c_int, _P, _P, _P, c_int64, _P]),
    "spc_plenoxel_decode": (c_int, [_P, c_int, c_int64, _P, c_int, _P, _P, c_int, c_float, c_float, _P, _P, _P]),
    "spc_plenoxel_decode_rows": (c_int, [_P, c_int, _P, c_int64, _P, c_int, _P, _P, c_int, c_float, c_float, _P, _P, _P]),
    "spc_plenoxel_crop_workspace": (c_int64, [c_int64]),
    "spc_plenoxel_crop_select": (c_int, [_P, c_int, _P, c_int64, _P, _P, _P, _P, _P, _P, _P, c_int64, _P]),
    "spc_seg_metrics": (c_int, [_P, _P, c_int64, c_int, c_int64, _P, _P]),
    "spc_seg_head_fwd": (c_int, [_P, c_int64, _P, _P, c_int64, c_int, c_int64, _P, _P, _P, _P, _P, _P]),
    "spc_inst_norm_fwd": (c_int, [_P, _P, c_int64, c_int, c_int, _P, _P, c_float, _P, _P, _P, _P, _P, _P]),
    "spc_inst_norm_bwd": (c_int, [_P, _P, _P, c_int64, c_int, c_int, _P, _P, _P, _P, _P, _P, _P]),
    "spc_inst_norm_force_scalar": (None, [c_int]),
    "spc_interp_corners": (c_int, [_P, c_int64, _P, _P, _P, _P]),
    "spc_interp_fwd": (c_int, [_P, _P, _P, c_int64, c_int, c_int, _P, _P]),
    "spc_interp_bwd": (c_int, [_P, _P, _P, c_int64, c_int64, c_int, c_int, _P, _P]),
    "spc_sgd_step": (c_int, [_P, _P, _P, c_int64, c_float, c_float, c_float, c_float, c_int, _P]),
}

_lib = None


class SparseConvLibraryError(RuntimeError):
    pass


def load():
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    path = Path(os.environ.get("SPARSECONV_B200_LIB", LIB_PATH))
    if not path.exists():
        raise SparseConvLibraryError(
            f"{path} is missing: build it with `python -m nerf_downstream_b200.build` "
            "(there is no CPU or PyTorch fallback for the sparse-convolution hot path)")
    lib = ctypes.CDLL(str(path))
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the header and the library disagree
        fn.restype = restype
        fn.argtypes = argtypes
    if lib.spc_abi_version() != 1:
        raise SparseConvLibraryError("ABI version mismatch")
    # measurement knobs of the convolution kernels (spc_debug_set; all 0 by default), e.g. "8=1" = no CTA-pair kernel
    for kv in filter(None, os.environ.get("SPARSECONV_B200_DEBUG_SET", "").split(",")):
        idx, val = kv.split("=")
        lib.spc_debug_set(int(idx), int(val))
    _lib = lib
    return lib


def last_error() -> str:
    return load().spc_last_error().decode("utf-8", "replace")


def check(rc: int, what: str) -> None:
    if rc != 0:
        raise RuntimeError(f"{what} failed: {last_error()}")


def ptr(t):
    """Device pointer of a CUDA tensor (None -> NULL).  Refuses host tensors."""
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("sparseconv_b200 kernels need CUDA tensors (no CPU fallback)")
    if not t.is_contiguous():
        raise RuntimeError("sparseconv_b200 kernels need contiguous tensors")
    if t.device.index != torch._C._cuda_getDevice():
        # launches go to the CURRENT device's current stream: a tensor on another device would be read through a
        # foreign pointer (illegal address, or an unordered peer access)
        raise RuntimeError(f"tensor lives on cuda:{t.device.index} but the current device is "
                           f"cuda:{torch._C._cuda_getDevice()}: call torch.cuda.set_device() for the device "
                           "the model and batch live on (one process per GPU)")
    return t.data_ptr()


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def stream():
    """cudaStream_t of torch's current stream on the current device (raw handle: this is called
    once per kernel launch, torch.cuda.current_stream() costs several microseconds of Python)."""
    if _raw_stream is not None:
        return _raw_stream(torch._C._cuda_getDevice())
    return torch.cuda.current_stream().cuda_stream


def launch_count() -> int:
    return int(load().spc_launch_count())
