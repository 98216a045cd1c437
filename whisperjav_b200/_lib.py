"""ctypes binding of libwjb200.so (include/wjb200.h).  Fails loudly when the library or a GPU
is missing -- there is no CPU fallback in the product path."""
from __future__ import annotations

import ctypes as C
import threading
from pathlib import Path

_HERE = Path(__file__).resolve().parent
LIB_PATH = _HERE / "libwjb200.so"


class WjbError(RuntimeError):
    pass


class Dims(C.Structure):
    _fields_ = [(k, C.c_int32) for k in (
        "n_mels", "n_audio_ctx", "n_audio_state", "n_audio_head", "n_audio_layer",
        "n_vocab", "n_text_ctx", "n_text_state", "n_text_head", "n_text_layer")]


class DecodeOpts(C.Structure):
    _fields_ = [(k, C.c_int32) for k in (
        "n_initial", "sot_index", "sample_len", "eot", "no_speech", "no_timestamps", "timestamp_begin",
        "suppress_blank", "blank_token", "apply_timestamp_rules", "max_initial_timestamp_index",
        "tokens_stride", "check_every")] + [("temperature", C.c_float), ("seed", C.c_uint32)]


class BeamBufs(C.Structure):
    """include/wjb200.h::wjb_beam_bufs"""
    _fields_ = [("n_audio", C.c_int32), ("beam_size", C.c_int32), ("max_candidates", C.c_int32), ("tokens", C.c_void_p),
                ("anc", C.c_void_p), ("sum_logprob", C.c_void_p), ("fin_tokens", C.c_void_p), ("fin_score", C.c_void_p),
                ("fin_len", C.c_void_p), ("fin_count", C.c_void_p), ("audio_done", C.c_void_p)]


_SIGS = {
    "wjb_abi_version": (C.c_int, []),
    "wjb_last_error": (C.c_char_p, []),
    "wjb_weights_bytes": (C.c_size_t, [C.POINTER(Dims)]),
    "wjb_weight_count": (C.c_int, [C.POINTER(Dims)]),
    "wjb_weight_info": (C.c_int, [C.POINTER(Dims), C.c_int, C.c_char_p, C.c_int, C.POINTER(C.c_size_t),
                                  C.POINTER(C.c_size_t), C.POINTER(C.c_int)]),
    "wjb_model_create": (C.c_int, [C.POINTER(Dims), C.c_void_p, C.POINTER(C.c_void_p)]),
    "wjb_model_destroy": (None, [C.c_void_p]),
    "wjb_logmel_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int]),
    "wjb_logmel_f16": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                                 C.c_int64, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "wjb_encoder_workspace_bytes": (C.c_size_t, [C.c_void_p, C.c_int]),
    "wjb_encoder_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "wjb_encoder_set_tap": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    "wjb_cross_kv_bytes": (C.c_size_t, [C.c_void_p, C.c_int]),
    "wjb_cross_kv": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "wjb_decode_workspace_bytes": (C.c_size_t, [C.c_void_p, C.c_int]),
    "wjb_decode_greedy": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(DecodeOpts), C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_int),
                                    C.c_void_p]),
    "wjb_decode_logits_stride": (C.c_int, [C.c_void_p]),
    "wjb_decode_set_trace": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]),
    "wjb_align_qk_bytes": (C.c_size_t, [C.c_void_p, C.c_int, C.c_int]),
    "wjb_decode_set_align": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "wjb_align_prefill_workspace_bytes": (C.c_size_t, [C.c_void_p, C.c_int, C.c_int]),
    "wjb_align_prefill": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_size_t, C.c_void_p]),
    "wjb_align_workspace_bytes": (C.c_size_t, [C.c_void_p, C.c_int, C.c_int]),
    "wjb_align_dtw": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "wjb_decode_beam": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(BeamBufs), C.POINTER(DecodeOpts), C.c_void_p, C.c_void_p, C.c_void_p,
                                  C.c_size_t, C.POINTER(C.c_int), C.c_void_p]),
    "wjb_gemm_f16": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int,
                               C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_void_p]),
    "wjb_gemm_splitk_workspace_bytes": (C.c_size_t, []),
    "wjb_gemm_f16_splitk": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p]),
    "wjb_gemm_step_f16": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "wjb_gemm_step_ln_f16": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                       C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "wjb_gemm_step_stats_f16": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                          C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "wjb_debug_gemm_trace": (None, [C.c_void_p]),
    "wjb_debug_set_pdl": (None, [C.c_int]),
    "wjb_debug_set_decode_flags": (None, [C.c_int]),
    "wjb_debug_get_decode_flags": (C.c_int, []),
    "wjb_layernorm_f16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "wjb_attention_encoder_f16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "wjb_attention_self_f16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "wjb_attention_cross_f16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "wjb_attention_cross_beam_f16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "wjb_profile_enable": (None, [C.c_int]),
    "wjb_profile_read": (C.c_int, [C.POINTER(C.c_float), C.POINTER(C.c_int), C.c_int]),
    "wjb_frame_head_f16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "wjb_vad_weights_bytes": (C.c_size_t, []),
    "wjb_vad_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int]),
    "wjb_vad_forward": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "wjb_scene_energy": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int64, C.c_void_p]),
}

EXPORTS = tuple(_SIGS)

_lock = threading.Lock()
_lib = None


def load() -> C.CDLL:
    """Load libwjb200.so (built in-tree by ``python -m whisperjav_b200.build``)."""
    global _lib
    with _lock:
        if _lib is None:
            if not LIB_PATH.exists():
                raise WjbError(f"{LIB_PATH} is missing: build it with `python -m whisperjav_b200.build` "
                               "(there is no CPU fallback)")
            lib = C.CDLL(str(LIB_PATH))
            for name, (res, args) in _SIGS.items():
                fn = getattr(lib, name)  # AttributeError if the export is missing
                fn.restype = res
                fn.argtypes = args
            if lib.wjb_abi_version() != 1:
                raise WjbError("libwjb200.so ABI version mismatch")
            _lib = lib
        return _lib


def check(status: int, what: str = "") -> None:
    if status != 0:
        msg = load().wjb_last_error()
        raise WjbError(f"{what}: {msg.decode() if msg else 'unknown error'}")


def make_dims(d) -> Dims:
    return Dims(*(int(getattr(d, k)) for k, _ in Dims._fields_))


def ptr(t) -> C.c_void_p:
    """Device/host pointer of a torch tensor (or None)."""
    return C.c_void_p(0 if t is None else t.data_ptr())


def stream_ptr() -> C.c_void_p:
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)
