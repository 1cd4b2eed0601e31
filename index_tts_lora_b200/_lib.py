"""ctypes binding of libbvg.so (include/bvg.h).  No fallback: if the library is missing or a
call fails, a RuntimeError is raised — the reference swallows its JIT-load failure into a torch
path (infer.py:381-388); north_star forbids that here."""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("BVG_LIB") or os.path.join(_HERE, "libbvg.so")   # BVG_LIB: experiment builds (tools/)

BVG_F32, BVG_BF16, BVG_F16, BVG_I16 = 0, 1, 2, 3
PREC_F32, PREC_BF16 = 0, 1

MAX_UPS, MAX_KERNELS, MAX_DIL = 8, 4, 3


class BvgConfig(C.Structure):
    _fields_ = [
        ("gpt_dim", C.c_int32),
        ("upsample_initial_channel", C.c_int32),
        ("num_upsamples", C.c_int32),
        ("upsample_rates", C.c_int32 * MAX_UPS),
        ("upsample_kernel_sizes", C.c_int32 * MAX_UPS),
        ("num_kernels", C.c_int32),
        ("resblock_kernel_sizes", C.c_int32 * MAX_KERNELS),
        ("resblock_dilation_sizes", (C.c_int32 * MAX_DIL) * MAX_KERNELS),
        ("speaker_embedding_dim", C.c_int32),
        ("cond_in_each_up_layer", C.c_int32),
        ("snake_logscale", C.c_int32),
    ]


class BvgTensorDesc(C.Structure):
    _fields_ = [
        ("name", C.c_char_p),
        ("data", C.c_void_p),
        ("dtype", C.c_int32),
        ("ndim", C.c_int32),
        ("shape", C.c_int64 * 4),
    ]


PROFILE_CLASSES = 4


class BvgProfile(C.Structure):
    _fields_ = [("ms", C.c_double * PROFILE_CLASSES), ("flops", C.c_double * PROFILE_CLASSES),
                ("bytes", C.c_double * PROFILE_CLASSES), ("launches", C.c_int32 * PROFILE_CLASSES)]


class BvgShardGeom(C.Structure):
    _fields_ = [("f_begin", C.c_int32), ("f_end", C.c_int32), ("f_total", C.c_int32),
                ("own_left", C.c_int32), ("own_right", C.c_int32), ("own_max", C.c_int32)]


class BvgEcapaTdnn(C.Structure):
    _fields_ = [("w", C.c_void_p), ("bias", C.c_void_p), ("bn_scale", C.c_void_p), ("bn_shift", C.c_void_p),
                ("cin", C.c_int32), ("cout", C.c_int32), ("k", C.c_int32), ("dil", C.c_int32), ("relu", C.c_int32)]


class BvgEcapaBlock(C.Structure):
    _fields_ = [("tdnn1", BvgEcapaTdnn), ("res2", BvgEcapaTdnn * 7), ("tdnn2", BvgEcapaTdnn),
                ("se_w1", C.c_void_p), ("se_b1", C.c_void_p), ("se_w2", C.c_void_p), ("se_b2", C.c_void_p)]


class BvgEcapaDesc(C.Structure):
    _fields_ = [("in_channels", C.c_int32), ("channels", C.c_int32), ("scale", C.c_int32), ("se_channels", C.c_int32),
                ("att_channels", C.c_int32), ("mfa_channels", C.c_int32), ("emb_dim", C.c_int32),
                ("block0", BvgEcapaTdnn), ("blocks", BvgEcapaBlock * 3),
                ("mfa", BvgEcapaTdnn), ("asp_tdnn", BvgEcapaTdnn), ("asp_conv", BvgEcapaTdnn),
                ("asp_ctx_w", C.c_void_p), ("asp_bn_scale", C.c_void_p), ("asp_bn_shift", C.c_void_p),
                ("fc_w", C.c_void_p), ("fc_b", C.c_void_p)]


# every symbol include/bvg.h declares: name -> (restype, argtypes)
_P, _I, _L = C.c_void_p, C.c_int, C.c_int64
SYMBOLS = {
    "bvg_version": (_I, []),
    "bvg_last_error": (C.c_char_p, []),
    "bvg_plan_create": (_I, [C.POINTER(BvgConfig), _I, C.POINTER(_P)]),
    "bvg_plan_destroy": (_I, [_P]),
    "bvg_plan_load_weights": (_I, [_P, C.POINTER(BvgTensorDesc), _I, _P]),
    "bvg_decode": (_I, [_P, _P, _I, C.POINTER(C.c_int32), _I, _I, _P, _P, _I, _I, _P]),
    "bvg_decode_host": (_I, [_P, _P, _I, C.POINTER(C.c_int32), _I, _I, _P, _P, _I, _I, _P]),
    "bvg_decode_shard": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _P, _P, _I, _I, _P]),
    "bvg_receptive_field_frames": (_I, [_P]),
    "bvg_shard_setup": (_I, [_P, C.POINTER(BvgShardGeom), _P]),
    "bvg_shard_halo_frames": (_I, [_P]),
    "bvg_shard_export": (_I, [_P, _P]),
    "bvg_shard_connect": (_I, [_P, _I, _P]),
    "bvg_shard_local_ptrs": (_I, [_P, C.POINTER(_P), C.POINTER(_P), C.POINTER(_P)]),
    "bvg_shard_connect_ptr": (_I, [_P, _I, _P, _P, _P]),
    "bvg_shard_run": (_I, [_P, _I, _P, _I, _P, _P, _I, _I, _I, _P]),
    "bvg_shard_error": (_I, [_P]),
    "bvg_shard_clear_error": (_I, [_P]),
    "bvg_plan_workspace_bytes": (_L, [_P]),
    "bvg_plan_last_launches": (_I, [_P]),
    "bvg_plan_set_profiling": (_I, [_P, _I]),
    "bvg_plan_read_profile": (_I, [_P, C.POINTER(BvgProfile)]),
    "bvg_set_tc_fir_max_channels": (_I, [_I]),
    "bvg_set_tc_narrow_max_channels": (_I, [_I]),
    "bvg_set_tc_split_min_channels": (_I, [_I]),
    "bvg_set_tc_residual_mma": (_I, [_I]),
    "bvg_set_pdl": (_I, [_I]),
    "bvg_set_graphs": (_I, [_I]),
    "bvg_set_tc_cluster": (_I, [_I]),
    "bvg_ecapa_create": (_I, [C.POINTER(BvgEcapaDesc), _I, C.POINTER(_P)]),
    "bvg_ecapa_destroy": (_I, [_P]),
    "bvg_ecapa_forward": (_I, [_P, _P, _I, _I, _I, _P, _P]),
    "bvg_activation1d": (_I, [_P, _P, _I, _I, _I, _I, _P, _P, _P, _P, _I, _P]),
    "bvg_amp_layer": (_I, [_P, _P, _P, _I, _I, _I, _I, _P, _P, _I, _I, _I, _P, _P, _P, _P, _I, _I, _P]),
    "bvg_conv_transpose1d": (_I, [_P, _P, _I, _I, _I, _I, _P, _P, _I, _I, _I, _P]),
}

_lib: Optional[C.CDLL] = None


def load() -> C.CDLL:
    """Load libbvg.so once; raise loudly if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"libbvg.so not found at {LIB_PATH}. Build it with "
            "`python -c 'import __graft_entry__ as g; g.build()'` (needs nvcc, sm_100a). "
            "There is no CPU or PyTorch fallback for the BigVGAN decode path.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status: int, what: str) -> None:
    if status != 0:
        msg = load().bvg_last_error()
        raise RuntimeError(f"{what} failed (status {status}): {msg.decode() if msg else '?'}")


def torch_dtype_code(dt) -> int:
    import torch

    return {torch.float32: BVG_F32, torch.bfloat16: BVG_BF16, torch.float16: BVG_F16,
            torch.int16: BVG_I16}[dt]


def stream_ptr(device) -> int:
    import torch

    return torch.cuda.current_stream(device).cuda_stream
