"""ctypes binding of libtim_b200.so (C ABI in include/tim_b200.h). No torch import here.

The library is the product: if it is missing or cannot be loaded this module raises — there is no
Python / PyTorch / CPU fallback for the TIM forward anywhere in tim_b200/.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libtim_b200.so")
ABI_VERSION = 2

TIM_OK = 0
STATUS_NAMES = {0: "TIM_OK", -1: "TIM_ERR_INVALID", -2: "TIM_ERR_CUDA", -3: "TIM_ERR_NO_DEVICE",
                -4: "TIM_ERR_WEIGHTS", -5: "TIM_ERR_NOMEM"}

MODALITY_CODES = {"audio_visual": 0, "visual": 1, "audio": 2}
VARIANT_CODES = {"recognition": 0, "detection": 1}


class tim_config(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "variant", "d_model", "nhead", "num_layers", "ff_dim", "vis_dim", "aud_dim", "num_feats",
        "input_modality", "data_modality", "include_verb_noun", "n_verb", "n_noun", "n_action", "n_audio",
        "compute_dtype")]


class tim_outputs(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("verb", "noun", "action", "audio", "reg_visual", "reg_audio", "feats")]


# every symbol include/tim_b200.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "tim_abi_version": (C.c_int, []),
    "tim_create": (C.c_int, [C.POINTER(C.c_void_p), C.POINTER(tim_config), C.c_int]),
    "tim_destroy": (None, [C.c_void_p]),
    "tim_last_error": (C.c_char_p, [C.c_void_p]),
    "tim_set_weight": (C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p, C.POINTER(C.c_int64), C.c_int, C.c_void_p]),
    "tim_weights_missing": (C.c_int, [C.c_void_p, C.c_char_p, C.c_size_t]),
    "tim_time_mlp_fwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "tim_encoder_fwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                  C.POINTER(tim_outputs), C.c_void_p]),
    "tim_forward_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                   C.POINTER(tim_outputs), C.c_int, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "tim_workspace_bytes": (C.c_size_t, [C.c_void_p]),
    "tim_launch_count": (C.c_uint64, [C.c_void_p]),
    "tim_seq_len": (C.c_int, [C.POINTER(tim_config), C.c_int, C.c_int]),
    "tim_profile_begin": (C.c_int, [C.c_void_p]),
    "tim_profile_end": (C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_uint64), C.c_int]),
    "tim_test_linear": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                  C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "tim_bench_linear": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float)]),
    "tim_fold_active": (C.c_int, [C.c_void_p]),
    "tim_encoder_fwd_indexed": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                          C.c_void_p]),
    "tim_label_queries": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float,
                                    C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "tim_smooth_labels": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int64, C.c_int, C.c_double, C.c_void_p, C.c_void_p]),
    "tim_nms_workspace_bytes": (C.c_size_t, [C.c_int64]),
    "tim_softnms_1d": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_float, C.c_float, C.c_int,
                                 C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "tim_nms_1d": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_float, C.c_int64,
                             C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "tim_det_decode": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_int, C.c_float, C.c_float, C.c_void_p,
                                 C.c_void_p, C.c_void_p]),
    "tim_det_count": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_float, C.c_void_p, C.c_void_p]),
    "tim_det_emit": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                               C.c_void_p, C.c_void_p]),
    "tim_host_chunk_schedule": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.c_int]),
    "tim_debug_role_prof": (C.c_int, [C.c_void_p]),
    "tim_bench_attention": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                      C.POINTER(C.c_float)]),
    "tim_test_attention": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                     C.c_void_p]),
    "tim_forward_host_ex": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                      C.POINTER(tim_outputs), C.c_int, C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_uint64),
                                      C.POINTER(C.c_uint64)]),
    "tim_fold_check": (C.c_int, [C.c_void_p]),
    "tim_index_check": (C.c_int, [C.c_void_p]),
    "tim_train_enable": (C.c_int, [C.c_void_p]),
    "tim_bind_grad": (C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p]),
    "tim_set_dropout": (C.c_int, [C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_uint64]),
    "tim_time_mlp_fwd_train": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "tim_time_mlp_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "tim_encoder_fwd_train": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                        C.POINTER(tim_outputs), C.c_void_p]),
    "tim_encoder_bwd": (C.c_int, [C.c_void_p, C.POINTER(tim_outputs), C.c_void_p, C.c_void_p]),
    "tim_train_tape_bytes": (C.c_size_t, [C.c_void_p]),
    "tim_comm_unique_id": (C.c_int, [C.c_void_p]),
    "tim_comm_init": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int]),
    "tim_allreduce_grads": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "tim_test_wgrad": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "tim_bench_wgrad": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                  C.POINTER(C.c_float)]),
    "tim_test_attention_bwd": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                         C.c_float, C.c_void_p]),
}

_lib = None


class tim_feature_bank(C.Structure):
    _fields_ = [("vis_bank", C.c_void_p), ("aud_bank", C.c_void_p), ("vis_rows", C.c_void_p), ("aud_rows", C.c_void_p),
                ("vis_bank_rows", C.c_int64), ("aud_bank_rows", C.c_int64), ("bank_dtype", C.c_int32)]


class TimError(RuntimeError):
    def __init__(self, status: int, msg: str):
        super().__init__(f"{STATUS_NAMES.get(status, status)}: {msg}")
        self.status = status


def load() -> C.CDLL:
    """Load the shared library (once). Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -m tim_b200.build` (nvcc, sm_100a). "
            "tim_b200 has no fallback path.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    v = lib.tim_abi_version()
    if v != ABI_VERSION:
        raise RuntimeError(f"libtim_b200 ABI version {v} != binding {ABI_VERSION}; rebuild the library")
    _lib = lib
    return lib


def check_nonneg(status: int, ctx=None) -> int:
    """for entry points that return a count / flag (>= 0) or a negative tim_status"""
    if status < 0:
        check(status, ctx)
    return status


def check(status: int, ctx=None) -> None:
    if status != TIM_OK:
        msg = load().tim_last_error(ctx)
        raise TimError(status, msg.decode() if msg else "")
