"""ctypes binding of the C ABI declared in include/audiocaption_b200.h.

There is NO CPU fallback: if the shared library is missing or a call fails this raises."""
import ctypes as C
import os

from .build import LIB_PATH

_lib = None

c_f32p = C.c_void_p
c_i64p = C.c_void_p

_SIGS = {
    "ac_version": (C.c_int, []),
    "ac_last_error": (C.c_char_p, []),
    "ac_launch_count": (C.c_int64, []),
    "ac_timing_enable": (None, [C.c_int]),
    "ac_timing_report": (C.c_int, [C.c_char_p, C.c_int]),
    "ac_frontend_create": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "ac_frontend_destroy": (None, [C.c_void_p]),
    "ac_frontend_num_frames": (C.c_int, [C.c_void_p, C.c_int]),
    "ac_logmel_fwd": (C.c_int, [C.c_void_p, c_f32p, C.c_int, C.c_int, c_f32p, c_f32p, C.c_void_p]),
    "ac_db_clamp": (C.c_int, [c_f32p, C.c_int64, c_f32p, C.c_float, C.c_void_p]),
    "ac_gemm": (C.c_int, [c_f32p, c_f32p, c_f32p, C.c_int, C.c_int, C.c_int, c_f32p, C.c_int, c_f32p, c_f32p, c_f32p,
                          C.c_int, C.c_int, C.c_void_p]),
    "ac_gemm_trace": (C.c_int, [C.c_int, C.c_void_p]),
    "ac_dwconv": (C.c_int, [c_f32p, c_f32p, c_f32p, c_f32p, c_f32p, c_f32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                            C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "ac_conv3x3": (C.c_int, [c_f32p, c_f32p, c_f32p, c_f32p, c_f32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                             C.c_void_p]),
    "ac_dwconv_partial_rows": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "ac_cnn14_num_tensors": (C.c_int, []),
    "ac_cnn14_out_dim": (C.c_int, []),
    "ac_cnn14_out_frames": (C.c_int, [C.c_int]),
    "ac_cnn14_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "ac_cnn14_create": (C.c_int, [C.POINTER(C.c_void_p), C.POINTER(C.c_int64), C.c_int, C.c_void_p, C.POINTER(C.c_void_p)]),
    "ac_cnn14_destroy": (None, [C.c_void_p]),
    "ac_cnn14_fwd": (C.c_int, [C.c_void_p, c_f32p, C.c_int, C.c_int, C.c_int, C.c_void_p, c_f32p, c_f32p, C.c_void_p,
                               C.c_size_t, C.c_void_p]),
    "ac_bigru_create": (C.c_int, [C.POINTER(C.c_void_p), C.POINTER(C.c_int64), C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                  C.POINTER(C.c_void_p)]),
    "ac_bigru_destroy": (None, [C.c_void_p]),
    "ac_bigru_out_dim": (C.c_int, [C.c_void_p]),
    "ac_bigru_workspace_bytes": (C.c_size_t, [C.c_void_p, C.c_int, C.c_int]),
    "ac_bigru_fwd": (C.c_int, [C.c_void_p, c_f32p, C.c_void_p, C.c_int, C.c_int, C.c_int, c_f32p, C.c_void_p, C.c_size_t,
                               C.c_void_p]),
    "ac_bah_num_tensors": (C.c_int, []),
    "ac_bah_create": (C.c_int, [C.POINTER(C.c_void_p), C.POINTER(C.c_int64), C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_void_p)]),
    "ac_bah_destroy": (None, [C.c_void_p]),
    "ac_bah_workspace_bytes": (C.c_size_t, [C.c_void_p, C.c_int, C.c_int]),
    "ac_bah_greedy": (C.c_int, [C.c_void_p, c_f32p, c_f32p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                C.c_void_p, c_f32p, c_f32p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "ac_bah_greedy_ex": (C.c_int, [C.c_void_p, c_f32p, c_f32p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                   c_f32p, c_i64p, C.c_void_p, c_f32p, c_f32p, c_f32p, c_f32p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "ac_bah_beam": (C.c_int, [C.c_void_p, c_f32p, c_f32p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float,
                              C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "ac_sed_num_tensors": (C.c_int, []),
    "ac_sed_segments": (C.c_int, [C.c_int]),
    "ac_sed_create": (C.c_int, [C.POINTER(C.c_void_p), C.POINTER(C.c_int64), C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_void_p)]),
    "ac_sed_destroy": (None, [C.c_void_p]),
    "ac_sed_workspace_bytes": (C.c_size_t, [C.c_void_p, C.c_int, C.c_int, C.c_int]),
    "ac_sed_fwd": (C.c_int, [C.c_void_p, c_f32p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, c_f32p, C.c_void_p,
                             C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "ac_effb2_create": (C.c_int, [C.POINTER(C.c_void_p), C.POINTER(C.c_int64), C.c_int, C.c_void_p, C.POINTER(C.c_void_p)]),
    "ac_effb2_destroy": (None, [C.c_void_p]),
    "ac_effb2_num_tensors": (C.c_int, []),
    "ac_effb2_out_frames": (C.c_int, [C.c_int]),
    "ac_effb2_out_dim": (C.c_int, []),
    "ac_effb2_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "ac_effb2_fwd": (C.c_int, [C.c_void_p, c_f32p, c_f32p, C.c_float, C.c_int, C.c_int, C.c_int, c_f32p,
                               C.c_void_p, C.c_size_t, C.c_void_p]),
    "ac_effb2_block_info": (C.c_int, [C.c_int, C.POINTER(C.c_int)]),
    "ac_masked_mean": (C.c_int, [c_f32p, c_i64p, C.c_int, C.c_int, C.c_int, c_f32p, C.c_void_p]),
    "ac_trm_create": (C.c_int, [C.POINTER(C.c_void_p), C.POINTER(C.c_int64), C.c_int, C.c_int, C.c_int, C.c_int,
                                C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_void_p)]),
    "ac_trm_destroy": (None, [C.c_void_p]),
    "ac_trm_num_tensors": (C.c_int, [C.c_int]),
    "ac_trm_workspace_bytes": (C.c_size_t, [C.c_void_p, C.c_int, C.c_int, C.c_int]),
    "ac_trm_greedy": (C.c_int, [C.c_void_p, c_f32p, c_i64p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                c_i64p, c_f32p, c_f32p, c_f32p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "ac_trm_trace": (C.c_int, [C.c_int, C.c_void_p]),
    "ac_trm_beam": (C.c_int, [C.c_void_p, c_f32p, c_i64p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float,
                              C.c_int, C.c_int, C.c_int, c_i64p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "ac_resample_out_len": (C.c_int, [C.c_int, C.c_int, C.c_int]),
    "ac_resample": (C.c_int, [c_f32p, C.c_int, C.c_int, c_f32p, C.c_int, C.c_int, C.c_int, C.c_int, c_f32p, C.c_void_p]),
    "ac_cnn14_set_precision": (C.c_int, [C.c_void_p, C.c_int]),
    "ac_cnn14_set_sm_limit": (C.c_int, [C.c_void_p, C.c_int]),
    "ac_sm_partition_create": (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    "ac_sm_partition_stream": (C.c_void_p, [C.c_void_p, C.c_int]),
    "ac_sm_partition_sms": (C.c_int, [C.c_void_p, C.c_int]),
    "ac_sm_partition_destroy": (None, [C.c_void_p]),
    "ac_sed_set_precision": (C.c_int, [C.c_void_p, C.c_int]),
    "ac_conv3x3_p": (C.c_int, [c_f32p, c_f32p, c_f32p, c_f32p, c_f32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                               C.c_int, C.c_void_p]),
    "ac_conv3x3_bf16": (C.c_int, [C.c_void_p, c_f32p, c_f32p, c_f32p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                  C.c_int, C.c_void_p]),
    # ---- training step
    "ac_cnn14_fwd_train": (C.c_int, [C.c_void_p, c_f32p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_float, C.c_float, C.c_uint64,
                                     c_f32p, c_f32p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "ac_trm_update": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.c_int, C.c_void_p]),
    "ac_trm_sample_forced": (C.c_int, [C.c_void_p, c_f32p, c_i64p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                       c_i64p, c_i64p, c_f32p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "ac_bigru_train_create": (C.c_int, [C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_int64), C.c_int, C.c_int,
                                        C.c_int, C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_void_p)]),
    "ac_bigru_train_destroy": (None, [C.c_void_p]),
    "ac_bigru_train_workspace_bytes": (C.c_size_t, [C.c_void_p, C.c_int, C.c_int]),
    "ac_bigru_train_refresh": (C.c_int, [C.c_void_p, C.c_void_p]),
    "ac_bigru_train_fwd": (C.c_int, [C.c_void_p, c_f32p, c_i64p, C.c_int, C.c_int, C.c_float, C.c_uint64, c_f32p, C.c_void_p,
                                     C.c_size_t, C.c_void_p]),
    "ac_bigru_train_bwd": (C.c_int, [C.c_void_p, c_f32p, c_i64p, c_f32p, C.c_int, C.c_int, C.c_float, C.c_uint64, c_f32p,
                                     C.c_void_p, C.c_size_t, C.c_void_p]),
    "ac_trm_train_num_tensors": (C.c_int, [C.c_int]),
    "ac_trm_train_create": (C.c_int, [C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_int64), C.c_int, C.c_int,
                                      C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_void_p)]),
    "ac_trm_train_destroy": (None, [C.c_void_p]),
    "ac_trm_train_workspace_bytes": (C.c_size_t, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]),
    "ac_trm_train_vocab_padded": (C.c_int, [C.c_void_p]),
    "ac_trm_train_refresh": (C.c_int, [C.c_void_p, C.c_void_p]),
    "ac_trm_train_memory_fwd": (C.c_int, [C.c_void_p, c_f32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_uint64,
                                          C.c_void_p, C.c_size_t, C.c_void_p]),
    "ac_trm_train_seq_fwd": (C.c_int, [C.c_void_p, c_i64p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, c_i64p, C.c_int,
                                       C.c_int, C.c_float, C.c_uint64, C.c_void_p, C.c_size_t, C.c_void_p]),
    "ac_trm_train_logits": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, c_f32p, c_f32p,
                                      C.c_int, C.c_void_p, C.c_size_t, C.c_void_p]),
    "ac_trm_train_bwd": (C.c_int, [C.c_void_p, c_f32p, C.c_void_p, C.c_int, c_i64p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                   c_f32p, c_i64p, C.c_int, C.c_int, C.c_float, C.c_uint64, c_f32p, C.c_void_p, C.c_size_t,
                                   C.c_void_p]),
    "ac_ls_ce_fwd_bwd": (C.c_int, [c_f32p, C.c_int, c_i64p, C.c_int, c_i64p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float,
                                   c_f32p, c_f32p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "ac_specaug_apply": (C.c_int, [c_f32p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p]),
    "ac_argmax_rows": (C.c_int, [c_f32p, C.c_int, C.c_int, C.c_int, c_i64p, c_f32p, C.c_void_p]),
    "ac_clip_adam_workspace_bytes": (C.c_size_t, []),
    "ac_clip_adam": (C.c_int, [c_f32p, c_f32p, c_f32p, c_f32p, C.c_int64, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float,
                               C.c_float, C.c_float, c_f32p, C.c_void_p, c_f32p, C.c_void_p, C.c_size_t, C.c_void_p]),
}

EXPORTED_SYMBOLS = tuple(_SIGS)


class AudioCaptionB200Error(RuntimeError):
    pass


def lib():
    """The loaded library; raises (loudly) when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise AudioCaptionB200Error(
                f"CUDA library {LIB_PATH} is missing -- run `python -m audiocaption_b200.build` "
                "(there is no CPU fallback)")
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(l, name)
            fn.restype, fn.argtypes = res, args
        _lib = l
    return _lib


def check(rc: int, what: str = ""):
    if rc != 0:
        raise AudioCaptionB200Error(f"{what} failed ({rc}): {lib().ac_last_error().decode()}")


def ptr(t):
    """device/host pointer of a torch tensor (None -> NULL)"""
    return None if t is None else C.c_void_p(t.data_ptr())


def pointer_table(tensors):
    """(void* array, int64 numel array, n) where a None entry becomes a NULL pointer (a frozen parameter's gradient)."""
    n = len(tensors)
    ptrs = (C.c_void_p * n)(*[None if t is None else t.data_ptr() for t in tensors])
    numels = (C.c_int64 * n)(*[0 if t is None else t.numel() for t in tensors])
    return ptrs, numels, n


def tensor_table(tensors):
    n = len(tensors)
    ptrs = (C.c_void_p * n)(*[t.data_ptr() for t in tensors])
    numels = (C.c_int64 * n)(*[t.numel() for t in tensors])
    return ptrs, numels, n


def current_stream():
    """Raw cudaStream_t of torch's current stream on the current device.  `torch.cuda.current_stream()` costs ~30 us per
    call (availability probes, Stream object); the raw getter is what it wraps (~1 us) -- a training step asks ~100 times."""
    import torch
    try:
        return C.c_void_p(torch._C._cuda_getCurrentRawStream(torch._C._cuda_getDevice()))
    except AttributeError:                  # private getters moved: the public (slow) path
        return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def timing_report():
    """{kernel_name: (launches, total_ms)} since timing was enabled; synchronises the device."""
    buf = C.create_string_buffer(1 << 16)
    check(lib().ac_timing_report(buf, len(buf)), "ac_timing_report")
    out = {}
    for line in buf.value.decode().splitlines():
        name, n, ms = line.split()
        out[name] = (int(n), float(ms))
    return out
