"""ctypes binding of the C ABI declared in include/trtllm_b200.h.

The product path has NO fallback: if the shared library is missing or a symbol cannot be
resolved this raises, and every wrapper raises on a non-zero return code."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libtrtllm_llama_b200.so")


class LibraryNotBuilt(RuntimeError):
    pass


class KernelError(RuntimeError):
    pass


vp, i32, i64, f32, sz = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_size_t

# name -> (restype, argtypes); mirrors include/trtllm_b200.h one to one
SIGNATURES = {
    "tb_version": (C.c_char_p, []),
    "tb_check_device": (i32, []),
    "tb_rmsnorm": (i32, [vp, vp, vp, vp, vp, f32, i32, i32, vp]),
    "tb_rmsnorm_quant": (i32, [vp, vp, vp, vp, vp, vp, vp, vp, f32, i32, i32, i32, i32, vp]),
    "tb_quantize_per_token": (i32, [vp, vp, vp, i32, i32, i32, vp]),
    "tb_quantize_tensor": (i32, [vp, vp, i64, vp, i32, vp]),
    "tb_gemv": (i32, [i32, vp, vp, vp, vp, vp, vp, vp, i32, i32, vp, i32, i32, i32, i32, vp]),
    "tb_gemv_max_rows": (i32, [i32, i32]),
    "tb_gemv_on_tensor_cores": (i32, [i32, i32, i32]),
    "tb_gemv_hint_next": (i32, [vp, sz, vp, sz]),
    "tb_gemv_fused": (i32, [i32, vp, vp, vp, vp, vp, vp, vp, i32, i32, vp, i32, i32, i32, i32, i32, vp, f32, vp]),
    "tb_gemm_tc_workspace_bytes": (sz, [i32, i32, i32]),
    "tb_gemm_tc_counter_bytes": (sz, []),
    "tb_gemm_tc": (i32, [i32, vp, i32, vp, vp, vp, vp, vp, i32, i32, vp, i32, i32, i32, vp, sz, vp, i32, i32, vp]),
    "tb_gemm_tc_swiglu": (i32, [i32, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, vp]),
    "tb_mmha_workspace_bytes": (sz, [i32, i32, i32]),
    "tb_mmha_num_splits": (i32, [i32, i32, i32, i32]),
    "tb_mmha_counter_bytes": (sz, [i32, i32]),
    "tb_mmha_decode": (i32, [vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, i32, f32, i32,
                             i32, vp]),
    "tb_mmha_set_mode": (i32, [i32]),
    "tb_mmha_decode_dev": (i32, [vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, i32, f32,
                                 i32, i32, vp]),
    "tb_mmha_decode_paged": (i32, [vp, vp, vp, i32, i32, vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, f32, i32,
                                   i32, vp]),
    "tb_context_attention_paged": (i32, [vp, vp, vp, i32, i32, vp, vp, vp, i32, i32, i32, i32, i32, f32, i32, vp]),
    "tb_context_attention_workspace_bytes": (sz, [i32, i32, i32]),
    "tb_context_attention": (i32, [vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, f32, i32, vp]),
    "tb_embedding": (i32, [vp, vp, vp, i32, i32, i32, vp]),
    "tb_swiglu": (i32, [vp, vp, vp, i32, i32, i32, vp]),
    "tb_swiglu_quant": (i32, [vp, vp, vp, vp, i32, i32, i32, vp]),
    "tb_add": (i32, [vp, vp, vp, i64, vp]),
    "tb_gather_last_token": (i32, [vp, vp, vp, i32, i32, i32, vp]),
    "tb_gather_last_token_packed": (i32, [vp, vp, vp, i32, i32, vp]),
    "tb_argmax": (i32, [vp, vp, i32, i32, i32, vp]),
    "tb_advance_step": (i32, [vp, vp, vp, vp, vp, i32, i32, vp]),
    "tb_half_to_float": (i32, [vp, vp, i64, vp]),
    "tb_fill_int": (i32, [vp, i32, i32, vp]),
    "tb_tile_int": (i32, [vp, i32, i32, vp]),
    "tb_force_ids": (i32, [vp, vp, vp, vp, i32, i32, vp]),
    "tb_unpack_rows": (i32, [vp, vp, vp, i32, i32, i32, vp]),
    "tb_pack_rows": (i32, [vp, vp, vp, i32, i32, i32, vp]),
    "tb_mmha_decode_beams": (i32, [vp, vp, vp, vp, i32, vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, i32, f32,
                                   i32, i32, vp]),
    "tb_beam_workspace_bytes": (sz, [i32, i32]),
    "tb_beam_init": (i32, [vp, vp, vp, vp, vp, vp, i32, i32, i32, vp]),
    "tb_beam_search_step": (i32, [vp, i32, i32, i32, i32, i32, f32, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, i32, vp, vp]),
    "tb_gather_tree": (i32, [vp, vp, vp, i32, i32, i32, i32, vp]),
    "tb_copy": (i32, [vp, vp, sz, vp]),
    "tb_gather_logits": (i32, [vp, vp, i32, i32, i32, vp]),
    "tb_sample": (i32, [vp, vp, i32, i32, i32, i32, f32, f32, C.c_uint64, vp, i32, vp, i32, vp, vp]),
    "tb_decode_step_max_batch": (i32, []),
    "tb_decode_step_create": (i32, [vp, vp, vp, vp]),
    "tb_decode_step_destroy": (None, [vp]),
    "tb_decode_step_launch": (i32, [vp, i32, vp]),
    "tb_decode_step_info": (i32, [vp, i32, vp, vp, vp]),
    "tb_decode_step_trace": (i32, [vp, i32, vp]),
    "tbrt_decode_step_handle": (vp, [vp]),
    "tb_mma_peak": (i32, [i32, i32, i32, vp, C.POINTER(C.c_double), vp]),
}



class TbrtConfig(C.Structure):
    """== tbrt_config (include/trtllm_b200_runtime.h)"""
    _fields_ = [("hidden", i32), ("heads", i32), ("inter", i32), ("layers", i32), ("vocab", i32), ("head_size", i32),
                ("rms_eps", f32), ("mode", i32), ("int8_kv", i32), ("max_batch", i32), ("max_input_len", i32),
                ("max_output_len", i32), ("tp_size", i32), ("tp_rank", i32), ("use_cuda_graph", i32),
                ("paged_kv_tokens_per_block", i32)]


class TbpField(C.Structure):
    """== tbp_field == nvinfer1::PluginField"""
    _fields_ = [("name", C.c_char_p), ("data", vp), ("type", i32), ("length", i32)]


class TbpDims(C.Structure):
    """== tbp_dims == nvinfer1::Dims"""
    _fields_ = [("nb_dims", i32), ("d", i32 * 8)]


class TbpTensorDesc(C.Structure):
    """== tbp_tensor_desc == nvinfer1::PluginTensorDesc"""
    _fields_ = [("dims", TbpDims), ("type", i32), ("format", i32), ("scale", f32)]


_P = C.POINTER
SIGNATURES.update({
    # include/trtllm_b200_plugin.h
    "tbp_init": (i32, [C.c_char_p]),
    "tbp_num_creators": (i32, []),
    "tbp_creator_name": (C.c_char_p, [i32]),
    "tbp_creator_fields": (i32, [C.c_char_p, _P(C.c_char_p), i32]),
    "tbp_create": (vp, [C.c_char_p, C.c_char_p, C.c_char_p, _P(TbpField), i32]),
    "tbp_deserialize": (vp, [C.c_char_p, C.c_char_p, C.c_char_p, vp, sz]),
    "tbp_clone": (vp, [vp]),
    "tbp_destroy": (None, [vp]),
    "tbp_type": (C.c_char_p, [vp]),
    "tbp_version": (C.c_char_p, [vp]),
    "tbp_namespace": (C.c_char_p, [vp]),
    "tbp_serialization_size": (sz, [vp]),
    "tbp_serialize": (i32, [vp, vp]),
    "tbp_nb_outputs": (i32, [vp]),
    "tbp_output_dims": (i32, [vp, i32, _P(TbpDims), i32, _P(TbpDims)]),
    "tbp_output_dtype": (i32, [vp, i32, _P(i32), i32]),
    "tbp_supports_format": (i32, [vp, i32, _P(TbpTensorDesc), i32, i32]),
    "tbp_workspace_size": (sz, [vp, _P(TbpTensorDesc), i32, _P(TbpTensorDesc), i32]),
    "tbp_initialize": (i32, [vp]),
    "tbp_enqueue": (i32, [vp, _P(TbpTensorDesc), _P(TbpTensorDesc), _P(vp), _P(vp), vp, vp]),
    "tb_comm_unique_id": (i32, [vp]),
    "tb_comm_init": (i32, [vp, _P(i32), i32, i32]),
    "initLibNvInferPlugins": (C.c_bool, [vp, C.c_char_p]),
    "getPluginRegistry": (vp, []),
    "getInferLibVersion": (i32, []),
    # include/trtllm_b200_runtime.h
    "tbrt_create": (vp, [_P(TbrtConfig)]),
    "tbrt_destroy": (None, [vp]),
    "tbrt_last_error": (C.c_char_p, []),
    "tbrt_set_tensor": (i32, [vp, C.c_char_p, vp, sz]),
    "tbrt_finalize": (i32, [vp]),
    "tbrt_device_bytes": (sz, [vp]),
    "tbrt_context": (i32, [vp, vp, vp, i32, i32, vp]),
    "tbrt_context_packed": (i32, [vp, vp, vp, i32, i32, i32, vp]),
    "tbrt_step": (i32, [vp, vp]),
    "tbrt_force_ids": (i32, [vp, vp, vp]),
    "tbrt_logits": (vp, [vp]),
    "tbrt_output_ids": (vp, [vp]),
    "tbrt_kv_cache": (vp, [vp, i32]),
    "tbrt_generate": (i32, [vp, vp, vp, i32, i32, i32, vp, vp]),
    "tbrt_last_launches": (i64, [vp]),
    "tbrt_set_end_id": (i32, [vp, i32]),
    "tbrt_set_decode_mode": (i32, [vp, i32]),
    "tbrt_kv_max_blocks_per_seq": (i32, [vp]),
    "tbrt_set_kv_blocks": (i32, [vp, vp, i32, i32, vp]),
    "tbrt_set_sampling": (i32, [vp, i32, f32, f32, C.c_uint64]),
    "tbrt_beam_begin": (i32, [vp, i32, f32, i32, vp]),
    "tbrt_beam_finalize": (i32, [vp, vp, vp, i32, vp]),
    "tbrt_fused_step_available": (i32, [vp]),
    "tbrt_last_steps": (i32, [vp]),
    "tb_finished": (i32, [vp, vp, i32, i32, i32, i32, i32, i32, vp]),
    "tbrt_ar_handle": (i32, [vp, vp]),
    "tbrt_ar_open": (i32, [vp, vp]),
    "tb_ar_create": (i32, [_P(vp), i32, i32, sz]),
    "tb_ar_create_ex": (i32, [_P(vp), i32, i32, sz, sz]),
    "tb_ar_extra": (vp, [vp, i32]),
    "tb_ar_extra_bytes": (sz, [vp]),
    "tb_decode_step_tp_bytes": (sz, [vp]),
    "tb_decode_step_tp_logits": (vp, [vp]),
    "tb_ar_destroy": (None, [vp]),
    "tb_ar_ipc_handle": (i32, [vp, vp]),
    "tb_ar_open_peers": (i32, [vp, vp]),
    "tb_ar_buffer": (vp, [vp, i32]),
    "tb_ar_allreduce": (i32, [vp, i32, vp, vp, i64, vp]),
})

_lib = None


def load_library(path: str | None = None):
    """dlopen the in-tree library and type every declared entry point."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise LibraryNotBuilt(
            f"{p} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU or PyTorch fallback for this path)")
    h = C.CDLL(p, mode=C.RTLD_GLOBAL)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(h, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = h
    return h


class _Lazy:
    def __getattr__(self, name):
        return getattr(load_library(), name)


lib = _Lazy()


def check(rc: int, what: str):
    if rc != 0:
        raise KernelError(f"{what} failed with code {rc}")
