"""ctypes binding of libkjarni_cuda.so (include/kjarni_cuda.h).

The library is the product; this module only declares its entry points.  It fails
loudly when the shared object is missing -- there is no CPU or PyTorch fallback.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libkjarni_cuda.so")

KJC_OK = 0
KJC_NULL_POINTER, KJC_INVALID_UTF8, KJC_MODEL_NOT_FOUND, KJC_LOAD_FAILED, KJC_INFERENCE_FAILED = 1, 2, 3, 4, 5
KJC_GPU_UNAVAILABLE, KJC_INVALID_CONFIG = 6, 7
STATUS_NAMES = {0: "Ok", 1: "NullPointer", 2: "InvalidUtf8", 3: "ModelNotFound", 4: "LoadFailed", 5: "InferenceFailed",
                6: "GpuUnavailable", 7: "InvalidConfig", 8: "Cancelled", 9: "Timeout", 10: "StreamEnded", 255: "Unknown"}

OUT_HIDDEN, OUT_POOLED, OUT_LOGITS = 0, 1, 2
POOL_MEAN, POOL_CLS, POOL_MAX, POOL_LAST = 0, 1, 2, 3
MASK_AUTO, MASK_ALLOC, MASK_NOALLOC = 0, 1, 2
SCAN_SEGMENT, SCAN_VECTORSTORE = 0, 1
ARCH_NAMES = {0: "bert", 1: "bert_prefixed", 2: "distilbert", 3: "roberta", 4: "mpnet"}
HEAD_NAMES = {0: None, 1: "dense_tanh", 2: "pre_relu", 3: "pooler_tanh", 4: "none"}
NO_ID = 0xFFFFFFFFFFFFFFFF
KERNEL_CLASSES = ("embed_ln", "gemm_qkv", "attention", "gemm_out", "layernorm", "gemm_ffn_up", "gemm_ffn_down", "output")


class KjcEncoderInfo(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "arch", "hidden_size", "num_layers", "num_heads", "intermediate_size", "vocab_size", "max_position_embeddings",
        "type_vocab_size", "position_offset", "head_kind", "num_labels", "device")] + [("layer_norm_eps", C.c_float)]


class KjcIndexDirInfo(C.Structure):
    _fields_ = [("dimension", C.c_int32), ("n_segments", C.c_int32), ("n_skipped", C.c_int32), ("reserved", C.c_int32),
                ("total_rows", C.c_uint64), ("max_docs_per_segment", C.c_uint64)]


class KjcForwardOptions(C.Structure):
    _fields_ = [("output", C.c_int32), ("pooling", C.c_int32), ("normalize", C.c_int32), ("mask_convention", C.c_int32)]


class KjarniCudaError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"{STATUS_NAMES.get(status, status)}: {message}")
        self.status = status
        self.message = message


_lib = None

# name -> (restype, argtypes); every symbol include/kjarni_cuda.h and include/kjarni_cuda_debug.h declare
_vp, _i, _u64, _f = C.c_void_p, C.c_int, C.c_uint64, C.c_float
SIGNATURES = {
    "kjc_last_error_message": (C.c_char_p, []),
    "kjc_clear_error": (None, []),
    "kjc_error_name": (C.c_char_p, [_i]),
    "kjc_version": (C.c_char_p, []),
    "kjc_device_count": (_i, []),
    "kjc_encoder_create": (_i, [C.c_char_p, _i, C.POINTER(_vp)]),
    "kjc_encoder_create_multi": (_i, [C.c_char_p, C.POINTER(_i), _i, C.POINTER(_vp)]),
    "kjc_encoder_device_count": (_i, [_vp]),
    "kjc_encoder_destroy": (None, [_vp]),
    "kjc_encoder_info": (_i, [_vp, C.POINTER(KjcEncoderInfo)]),
    "kjc_encoder_label": (C.c_char_p, [_vp, _i]),
    "kjc_encoder_forward": (_i, [_vp, _vp, _vp, _vp, _i, _i, C.POINTER(KjcForwardOptions), _vp]),
    "kjc_encoder_forward_device_async": (_i, [_vp, _vp, _vp, _vp, _i, _i, C.POINTER(KjcForwardOptions), _vp, _vp]),
    "kjc_encoder_micro_batch": (_i, [_vp, _i]),
    "kjc_encoder_chained": (_i, [_vp]),
    "kjc_encoder_set_fp32_residual": (_i, [_vp, _i]),
    "kjc_encoder_last_launch_count": (C.c_int64, [_vp]),
    "kjc_encoder_set_profiling": (_i, [_vp, _i]),
    "kjc_encoder_get_profile": (_i, [_vp, _vp, _vp]),
    "kjc_softmax_rows": (None, [_vp, _i, _i]),
    "kjc_sigmoid_rows": (None, [_vp, _i, _i]),
    "kjc_tokenizer_create": (_i, [C.c_char_p, _i, C.POINTER(_vp)]),
    "kjc_tokenizer_destroy": (None, [_vp]),
    "kjc_tokenizer_token_to_id": (_i, [_vp, C.c_char_p, C.POINTER(C.c_uint32)]),
    "kjc_tokenizer_encode_batch": (_i, [_vp, _vp, _vp, _i, _i, _vp, _vp, _vp, _i, C.POINTER(_i)]),
    "kjc_index_create": (_i, [_i, _u64, _u64, _i, C.POINTER(_vp)]),
    "kjc_index_destroy": (None, [_vp]),
    "kjc_index_len": (_u64, [_vp]),
    "kjc_index_dim": (_i, [_vp]),
    "kjc_index_dir_info": (_i, [C.c_char_p, C.POINTER(KjcIndexDirInfo)]),
    "kjc_index_dir_segment_lens": (_i, [C.c_char_p, _vp, _i]),
    "kjc_index_part_range": (_i, [_u64, _i, _i, C.POINTER(_u64), C.POINTER(_u64)]),
    "kjc_index_open_dir": (_i, [C.c_char_p, _i, _i, _i, C.POINTER(_vp)]),
    "kjc_index_id_base": (_u64, [_vp]),
    "kjc_index_add_rows": (_i, [_vp, _vp, _u64]),
    "kjc_index_load_vectors_bin": (_i, [_vp, C.c_char_p]),
    "kjc_index_append_synthetic": (_i, [_vp, C.c_uint32, _u64, _u64]),
    "kjc_index_get_rows": (_i, [_vp, _u64, _u64, _vp]),
    "kjc_index_search": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp, _vp]),
    "kjc_index_search_device_async": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "kjc_index_search_device": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "kjc_topk_merge_device_async": (_i, [_i, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "kjc_packed_record_bytes": (C.c_size_t, [_i, _i]),
    "kjc_topk_merge_packed_device_async": (_i, [_i, _vp, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "kjc_index_last_launch_count": (C.c_int64, [_vp]),
    "kjc_sharded_index_create": (_i, [_i, _u64, C.POINTER(_i), _i, C.POINTER(_vp)]),
    "kjc_sharded_index_open_dir": (_i, [C.c_char_p, C.POINTER(_i), _i, C.POINTER(_vp)]),
    "kjc_sharded_index_destroy": (None, [_vp]),
    "kjc_sharded_index_len": (_u64, [_vp]),
    "kjc_sharded_index_dim": (_i, [_vp]),
    "kjc_sharded_index_shards": (_i, [_vp]),
    "kjc_sharded_index_shard_len": (_u64, [_vp, _i]),
    "kjc_sharded_index_add_rows": (_i, [_vp, _vp, _u64]),
    "kjc_sharded_index_append_synthetic": (_i, [_vp, C.c_uint32, _u64]),
    "kjc_sharded_index_search": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp, _vp]),
    "kjc_sharded_index_last_launch_count": (C.c_int64, [_vp]),
    "kjc_index_unverified_count": (C.c_int64, [_vp]),
    "kjc_dbg_index_set_filter": (_i, [_vp, _f, _i]),
    "kjc_cosine_similarity": (_f, [_vp, _vp, C.c_size_t]),
    "kjc_dbg_gemm": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "kjc_dbg_gemm_ln_gemm": (_i, [_vp, _vp, _vp, _vp, _vp, _f, _vp, _i, _i, _vp, _vp, _i, _i, _i, _vp, _vp, _i, _vp]),
    "kjc_dbg_gemm_ln": (_i, [_vp, _vp, _vp, _vp, _vp, _f, _vp, _i, _i, _vp, _i, _vp]),
    "kjc_dbg_gemm_ln_h": (_i, [_vp, _vp, _vp, _vp, _vp, _f, _vp, _i, _i, _i, _vp, _i, _vp]),
    "kjc_dbg_experimental_kernels": (_i, []),
    "kjc_dbg_ffn_ln": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _f, _i, _i, _i, _vp, _i, _vp]),
    "kjc_dbg_gemm_time": (_i, [_i, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "kjc_dbg_attention": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "kjc_dbg_encoder_head": (_i, [_vp, _vp, _i, _i, _vp]),
}


def lib() -> C.CDLL:
    """Loads libkjarni_cuda.so; raises if it has not been built (`python -c 'import __graft_entry__ as g; g.build()'`)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: build it with `make -C kjarni_b200/csrc` "
                              "(or __graft_entry__.build()); kjarni_b200 has no CPU fallback")
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(l, name)  # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def check(status: int) -> None:
    if status != KJC_OK:
        msg = lib().kjc_last_error_message()
        raise KjarniCudaError(status, msg.decode("utf-8", "replace") if msg else "")
