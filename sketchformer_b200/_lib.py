"""ctypes binding of libskf.so (include/skf.h).  The CUDA library is the product path: importing
this module never falls back to a CPU implementation — a missing or unloadable library raises."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SKF_LIB") or os.path.join(HERE, "libskf.so")     # SKF_LIB: a timing-experiment build (tools/)


class SkfError(RuntimeError):
    pass


class SkfCfg(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("B", "T", "d", "f", "H", "L", "V", "C", "u", "continuous",
                                         "blind_decoder_mask", "do_classification", "do_reconstruction",
                                         "warmup_steps")] + \
               [(n, C.c_float) for n in ("class_weight", "recon_weight", "adam_beta1", "adam_beta2", "adam_eps")] + \
               [(n, C.c_int32) for n in ("world_size", "gemm_backend", "attn_backend", "gemm_precision")] + \
               [("dropout_rate", C.c_float), ("dropout_seed", C.c_uint32), ("attn_version", C.c_int32),
                ("class_buffer_layers", C.c_int32), ("class_dropout", C.c_float)]


class SkfParamEntry(C.Structure):
    _fields_ = [("name", C.c_char * 56), ("offset", C.c_int64), ("rows", C.c_int32), ("cols", C.c_int32)]


class SkfGemmEpilogue(C.Structure):
    _fields_ = [("bias", C.c_void_p), ("resid", C.c_void_p), ("ld_resid", C.c_int64), ("mask", C.c_void_p),
                ("ld_mask", C.c_int64), ("act", C.c_int32), ("accumulate", C.c_int32), ("round_tf32", C.c_int32),
                ("split_k", C.c_int32), ("a_lo", C.c_void_p), ("b_lo", C.c_void_p), ("colsum", C.c_void_p),
                ("b_hi16", C.c_void_p), ("b_lo16", C.c_void_p)]


P, I, L, F, SZ = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_size_t

# name -> (restype, argtypes); every symbol declared in include/skf.h
SIGNATURES = {
    "skf_last_error": (C.c_char_p, []),
    "skf_set_option": (I, [C.c_char_p, I]),
    "skf_version": (I, []),
    "skf_abi_sizeof": (SZ, [I]),
    "skf_device_ok": (I, []),
    "skf_param_count": (L, [C.POINTER(SkfCfg)]),
    "skf_param_table": (I, [C.POINTER(SkfCfg), C.POINTER(SkfParamEntry), I]),
    "skf_embed_fwd": (I, [P, L, P, P, P, I, I, I, I, I, P]),
    "skf_dropout": (I, [P, P, L, F, C.c_uint32, C.c_uint32, P, P]),
    "skf_embed_bwd": (I, [P, L, P, P, I, I, I, I, P]),
    "skf_positional_encoding": (I, [P, I, I, P]),
    "skf_pad_flags": (I, [P, L, P, I, I, P]),
    "skf_pad_flags_cont": (I, [P, L, P, I, I, P]),
    "skf_embed_cont_fwd": (I, [P, L, P, P, P, P, I, I, I, I, P]),
    "skf_embed_cont_bwd": (I, [P, L, P, P, P, I, I, I, P]),
    "skf_cont_loss": (I, [P, L, P, L, I, I, F, P, L, P, P]),
    "skf_gemm": (I, [P, L, L, P, L, L, P, L, I, I, I, C.POINTER(SkfGemmEpilogue), I, P]),
    "skf_ffn_fwd": (I, [P, P, P, P, P, P, P, P, I, I, I, P, P, P, P, P]),
    "skf_ffn_bwd": (I, [P, P, P, P, P, P, P, P, P, P, P, P, I, I, I, P, P, P, P, P]),
    "skf_logits_ce_fwd": (I, [P, P, P, P, L, I, I, I, I, F, P, L, P, P, P, P, P]),
    "skf_logits_ce_bwd": (I, [P, P, L, P, P, P, P, I, I, I, I, P, P, P]),
    "skf_split_lo": (I, [P, P, L, P]),
    "skf_split_bf16": (I, [P, P, P, L, P]),
    "skf_attn_fwd": (I, [P, L, P, L, P, L, P, L, P, P, I, I, I, I, I, I, I, I, P]),
    "skf_attn_bwd": (I, [P, L, P, L, P, L, P, L, P, L, P, P, P, L, P, L, P, L, P, I, I, I, I, I, I, I, I, P]),
    "skf_attn_image_bytes": (SZ, [I, I, I, I, I]),
    "skf_attn_fwd_img": (I, [P, L, P, L, P, L, P, L, P, P, I, I, I, I, I, I, P, P]),
    "skf_attn_bwd_img": (I, [P, L, P, L, P, L, P, L, P, L, P, P, P, L, P, L, P, L, P, I, I, I, I, I, I, P, P, P]),
    "skf_add_layernorm_fwd": (I, [P, P, P, P, P, P, P, P, P, I, I, I, P]),
    "skf_layernorm_bwd": (I, [P, P, P, P, P, P, P, P, I, I, I, P]),
    "skf_add_dropout_layernorm_fwd": (I, [P, P, P, P, P, P, P, P, P, I, I, F, C.c_uint32, C.c_uint32, P, P]),
    "skf_layernorm_dropout_bwd": (I, [P, P, P, P, P, P, P, P, P, I, I, F, C.c_uint32, C.c_uint32, P, P]),
    "skf_colsum": (I, [P, L, P, I, I, P]),
    "skf_pool_fwd": (I, [P, P, P, P, P, I, I, I, I, I, P]),
    "skf_pool_bwd": (I, [P, P, P, P, P, P, I, P, P, I, I, I, I, I, P]),
    "skf_expand_fwd": (I, [P, P, P, P, I, I, I, I, P]),
    "skf_expand_bwd": (I, [P, P, P, P, I, P, P, I, I, I, P]),
    "skf_kv_expand_fwd": (I, [P, P, P, P, P, P, I, I, I, P]),
    "skf_kv_expand_bwd": (I, [P, P, P, P, P, P, P, P, P, P, I, I, I, P]),
    "skf_xattn_fold_fwd": (I, [P, L, P, P, P, P, P, P, P, L, I, I, I, I, I, P]),
    "skf_xattn_fold_bwd": (I, [P, L, P, P, P, P, P, P, P, L, P, L, P, P, P, P, P, I, I, I, I, I, P]),
    "skf_add_row_broadcast": (I, [P, P, I, I, P]),
    "skf_token_ce": (I, [P, L, P, L, I, I, I, F, P, L, P, P, I, P]),
    "skf_class_ce": (I, [P, L, P, I, I, F, P, L, P, P, P, P]),
    "skf_adam_step": (I, [P, P, P, P, L, P, I, I, F, F, F, F, P]),
    "skf_workspace_bytes": (SZ, [C.POINTER(SkfCfg)]),
    "skf_plan_create": (I, [C.POINTER(SkfCfg), P, P, P, P, P, SZ, C.POINTER(P)]),
    "skf_plan_destroy": (None, [P]),
    "skf_plan_init": (I, [P, P]),
    "skf_plan_tensor": (I, [P, C.c_char_p, C.POINTER(SZ), C.POINTER(SZ)]),
    "skf_train_step": (I, [P, P, P, I, P]),
    "skf_apply_adam": (I, [P, P]),
    "skf_encode": (I, [P, P, P]),
    "skf_forward": (I, [P, P, P]),
    "skf_decode_begin": (I, [P, P, P, L, L, P]),
    "skf_decode_steps": (I, [P, I, I, P]),
    "skf_tokenize_strokes": (I, [P, P, I, P, I, F, P, I, P]),
    "skf_plan_launches": (L, [P]),
    "skf_debug_trace": (I, [P, I]),
    "skf_nccl_version": (I, []),
    "skf_comm_unique_id": (I, [P]),
    "skf_comm_create": (I, [P, I, I, C.POINTER(P)]),
    "skf_comm_destroy": (I, [P]),
    "skf_allreduce_bucket": (I, [P, L, P, P]),
    "skf_grad_buckets": (I, [C.POINTER(SkfCfg), I, C.POINTER(L), C.POINTER(L)]),
    "skf_plan_set_comm": (I, [P, P, I]),
}

_lib = None


def load():
    """dlopen libskf.so and bind every declared symbol.  Raises SkfError if it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SkfError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                       "(there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)   # AttributeError if the header and the library disagree
        fn.restype = res
        fn.argtypes = args
    for which, struct in enumerate((SkfCfg, SkfGemmEpilogue, SkfParamEntry)):
        if lib.skf_abi_sizeof(which) != C.sizeof(struct):
            raise SkfError(f"{struct.__name__}: ctypes layout ({C.sizeof(struct)} bytes) does not match libskf.so "
                           f"({lib.skf_abi_sizeof(which)} bytes): rebuild the library or update _lib.py")
    # A/B switches for measurement runs: SKF_OPTIONS="pdl=0,attn3=3" -> skf_set_option(name, value) once at load
    for clause in filter(None, os.environ.get("SKF_OPTIONS", "").split(",")):
        name, _, value = clause.partition("=")
        if lib.skf_set_option(name.strip().encode(), int(value)) != 0:
            raise SkfError(f"SKF_OPTIONS: unknown option {name!r}")
    _lib = lib
    return lib


def check(status, what=""):
    if status != 0:
        msg = load().skf_last_error()
        raise SkfError(f"{what} failed ({status}): {msg.decode() if msg else ''}")
