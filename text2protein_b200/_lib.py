"""ctypes binding of libt2p.so (C ABI declared in include/t2p.h).

There is no CPU fallback: if the library is missing or fails to load, every entry into the native path raises.
Build it with ``python -c "import __graft_entry__ as g; g.build()"`` (or ``python -m text2protein_b200._build``).
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libt2p.so")

F32, BF16, F64, I64, U8 = 0, 1, 2, 3, 4


class UnetCfg(C.Structure):
    _fields_ = [("num_channels", C.c_int32), ("max_res_num", C.c_int32), ("nf", C.c_int32),
                ("n_ch_mult", C.c_int32), ("ch_mult", C.c_int32 * 16), ("num_res_blocks", C.c_int32),
                ("n_attn_resolutions", C.c_int32), ("attn_resolutions", C.c_int32 * 16), ("n_heads", C.c_int32),
                ("context_dim", C.c_int32), ("num_scales", C.c_int32), ("scale_by_sigma", C.c_int32),
                ("compute_dtype", C.c_int32)]


class StepArgs(C.Structure):
    _fields_ = [("x", C.c_void_p), ("score", C.c_void_p), ("score_dtype", C.c_int32), ("score_nhwc", C.c_int32),
                ("sigmas", C.c_void_p), ("labels", C.c_void_p), ("G", C.c_void_p), ("sqrt_alpha", C.c_void_p),
                ("alpha", C.c_void_p), ("probability_flow", C.c_int32), ("snr", C.c_float), ("mask", C.c_void_p),
                ("x_init", C.c_void_p), ("x_mean_out", C.c_void_p), ("seed", C.c_uint64), ("stream_id", C.c_int64),
                ("sample_offset", C.c_int64), ("B", C.c_int32), ("C", C.c_int32), ("HW", C.c_int32),
                ("workspace", C.c_void_p), ("conditioned_in_place", C.c_int32), ("symmetrize", C.c_int32),
                ("x_out", C.c_void_p), ("W", C.c_int32), ("reserved", C.c_int32)]


class RunArgs(C.Structure):
    _fields_ = [("x", C.c_void_p), ("x_mean", C.c_void_p), ("mask", C.c_void_p), ("x_init", C.c_void_p),
                ("label_table", C.c_void_p), ("g_table", C.c_void_p), ("num_iters", C.c_int32),
                ("n_steps", C.c_int32), ("snr", C.c_float), ("probability_flow", C.c_int32), ("seed", C.c_uint64),
                ("sample_offset", C.c_int64), ("B", C.c_int32), ("use_graph", C.c_int32), ("symmetrize", C.c_int32),
                ("reserved", C.c_int32), ("peers", C.c_void_p)]


class ConvArgs(C.Structure):
    _fields_ = [("a0", C.c_void_p), ("c0", C.c_int32), ("a1", C.c_void_p), ("c1", C.c_int32), ("B", C.c_int32),
                ("H", C.c_int32), ("W", C.c_int32), ("ksize", C.c_int32), ("w", C.c_void_p), ("N", C.c_int32),
                ("bias", C.c_void_p), ("rowbias", C.c_void_p), ("rowbias_ld", C.c_int32), ("residual", C.c_void_p),
                ("res_up", C.c_int32), ("alpha", C.c_float), ("out", C.c_void_p), ("out_dtype", C.c_int32),
                ("in_dtype", C.c_int32), ("stat_part", C.c_void_p), ("x0", C.c_void_p), ("xc0", C.c_int32),
                ("x1", C.c_void_p), ("xc1", C.c_int32), ("gn_scale", C.c_void_p), ("gn_shift", C.c_void_p),
                ("gno_gamma", C.c_void_p), ("gno_beta", C.c_void_p), ("gno_groups", C.c_int32), ("gno_eps", C.c_float)]


class GemmRecord(C.Structure):
    _fields_ = [("M", C.c_int64), ("N", C.c_int32), ("K", C.c_int32), ("ksize", C.c_int32),
                ("tensor_core", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("ms", C.c_float)]


ABI_VERSION = 3
IPC_HANDLE_BYTES = 64
# `which` codes of t2p_sizeof / t2p_struct_layout
STRUCTS = {0: UnetCfg, 1: StepArgs, 2: RunArgs, 3: ConvArgs, 4: GemmRecord}

# name -> (restype, argtypes); must list every symbol include/t2p.h declares (tests/test_abi.py checks)
SIGNATURES = {
    "t2p_last_error": (C.c_char_p, []),
    "t2p_abi_version": (C.c_int, []),
    "t2p_sizeof": (C.c_int, [C.c_int]),
    "t2p_struct_layout": (C.c_int, [C.c_int, C.POINTER(C.c_int32), C.c_int]),
    "t2p_unet_forward_t": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                     C.c_void_p]),
    "t2p_peer_mailbox_create": (C.c_int, [C.c_int, C.POINTER(C.c_void_p), C.c_void_p]),
    "t2p_peer_group_open": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int64, C.POINTER(C.c_void_p)]),
    "t2p_peer_group_close": (None, [C.c_void_p]),
    "t2p_unet_create": (C.c_int, [C.POINTER(UnetCfg), C.POINTER(C.c_void_p)]),
    "t2p_unet_destroy": (None, [C.c_void_p]),
    "t2p_unet_num_params": (C.c_int, [C.c_void_p]),
    "t2p_unet_param_info": (C.c_int, [C.c_void_p, C.c_int, C.c_char_p, C.c_int, C.POINTER(C.c_int64),
                                      C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "t2p_unet_load": (C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p, C.POINTER(C.c_int64), C.c_int, C.c_int,
                                C.c_void_p]),
    "t2p_unet_finalize": (C.c_int, [C.c_void_p, C.c_void_p]),
    "t2p_unet_set_context": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "t2p_unet_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "t2p_unet_set_debug": (C.c_int, [C.c_void_p, C.c_int]),
    "t2p_unet_set_fused_groupnorm": (C.c_int, [C.c_void_p, C.c_int]),
    "t2p_unet_set_epilogue_groupnorm": (C.c_int, [C.c_void_p, C.c_int]),
    "t2p_unet_tap": (C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int64, C.POINTER(C.c_int64), C.c_void_p]),
    "t2p_unet_set_profile": (C.c_int, [C.c_void_p, C.c_int]),
    "t2p_unet_profile_read": (C.c_int, [C.c_void_p, C.POINTER(GemmRecord), C.c_int]),
    "t2p_unet_workspace_bytes": (C.c_int64, [C.c_void_p]),
    "t2p_unet_launches_per_forward": (C.c_int64, [C.c_void_p]),
    "t2p_corrector_workspace_bytes": (C.c_int64, [C.c_int, C.c_int64]),
    "t2p_predictor_step": (C.c_int, [C.POINTER(StepArgs), C.c_void_p]),
    "t2p_corrector_step": (C.c_int, [C.POINTER(StepArgs), C.c_void_p]),
    "t2p_philox_normal": (C.c_int, [C.c_uint64, C.c_int64, C.c_int64, C.c_int64, C.c_float, C.c_void_p,
                                    C.c_void_p]),
    "t2p_philox_bits": (C.c_int, [C.c_uint64, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p]),
    "t2p_pc_run": (C.c_int, [C.c_void_p, C.POINTER(RunArgs), C.c_void_p]),
    "t2p_unet_set_context_tokens": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_void_p, C.c_int, C.c_int,
                                              C.c_void_p]),
    "t2p_embed_tokens": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_int, C.c_void_p, C.c_int64, C.c_void_p,
                                   C.c_void_p]),
    "t2p_length_mask": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "t2p_inpaint_mask": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "t2p_condition_mask": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                     C.c_void_p, C.c_void_p]),
    "t2p_postprocess_6d": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "t2p_conv2d": (C.c_int, [C.POINTER(ConvArgs), C.c_void_p]),
    "t2p_conv2d_stat_tile": (C.c_int, [C.POINTER(ConvArgs)]),
    "t2p_conv2d_fuses_groupnorm": (C.c_int, [C.POINTER(ConvArgs)]),
    "t2p_conv2d_normalises_output": (C.c_int, [C.POINTER(ConvArgs)]),
    "t2p_final_conv": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                 C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "t2p_groupnorm": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                C.c_int, C.c_float, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                C.c_void_p]),
    "t2p_groupnorm_small": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                      C.c_int, C.c_float, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "t2p_groupnorm_apply": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                      C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "t2p_layernorm": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_float, C.c_int,
                                C.c_void_p, C.c_void_p]),
    "t2p_geglu": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "t2p_attention": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                C.c_int, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_float, C.c_int, C.c_int,
                                C.c_void_p]),
}

_lib = None


class NativeError(RuntimeError):
    pass


def check_layout(handle):
    """sizeof and every field offset of the ctypes mirrors against the structs the library was compiled with."""
    for which, cls in STRUCTS.items():
        if handle.t2p_sizeof(which) != C.sizeof(cls):
            raise NativeError(f"{cls.__name__}: sizeof {C.sizeof(cls)} here, {handle.t2p_sizeof(which)} in libt2p.so")
        n = len(cls._fields_)
        offs = (C.c_int32 * n)()
        if handle.t2p_struct_layout(which, offs, n) != n:
            raise NativeError(f"{cls.__name__}: field count differs from libt2p.so")
        for (fname, _), off in zip(cls._fields_, offs):
            if getattr(cls, fname).offset != off:
                raise NativeError(f"{cls.__name__}.{fname}: offset {getattr(cls, fname).offset} here, {off} in libt2p.so")


def use_library(path):
    """Selects another build of the library (e.g. libt2p_knobs.so, the -DT2P_TIMING_KNOBS build that tools/ and
    ``bench.py --lib knobs`` use) -- before the first native call."""
    global LIB_PATH
    if _lib is not None:
        raise NativeError("the native library is already loaded")
    LIB_PATH = path if os.path.isabs(path) else os.path.join(_HERE, path)


def lib():
    """Loads libt2p.so once; raises (never falls back) when it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise NativeError(f"{LIB_PATH} is missing: build it with __graft_entry__.build(); "
                              "text2protein_b200 has no CPU or PyTorch fallback for the sampling path")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        if handle.t2p_abi_version() != ABI_VERSION:
            raise NativeError("libt2p.so ABI version mismatch; rebuild")
        check_layout(handle)
        _lib = handle
    return _lib


def check(rc):
    if rc != 0:
        raise NativeError(lib().t2p_last_error().decode("utf-8", "replace"))


def current_stream():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def device_of(t):
    """Context manager making ``t``'s device current for the native calls inside it: the library launches on the
    current device (its per-device kernel attributes and SM counts are keyed by cudaGetDevice)."""
    import torch
    return torch.cuda.device(t.device)


def ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def torch_dtype_code(dt):
    import torch
    return {torch.float32: F32, torch.bfloat16: BF16, torch.float64: F64, torch.int64: I64, torch.uint8: U8,
            torch.bool: U8}[dt]
