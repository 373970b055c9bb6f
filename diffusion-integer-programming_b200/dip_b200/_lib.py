"""ctypes binding of libdip_b200.so (C-ABI declared in include/dip_b200.h).

The library is built in-tree by ``__graft_entry__.build()`` / ``dip_b200.build.build()``; there is
no fallback of any kind: if the shared object is missing, or a call fails, an exception is raised.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
PKG_ROOT = os.path.dirname(_HERE)
LIB_PATH = os.path.join(PKG_ROOT, "lib", "libdip_b200.so")

c_float_p = C.POINTER(C.c_float)
vp = C.c_void_p


class DipError(RuntimeError):
    pass


class Rows(C.Structure):
    _fields_ = [("shared", vp), ("batch", vp), ("T", C.c_int), ("split", C.c_int), ("shared_stride", C.c_int64)]


class GemmArgs(C.Structure):
    _fields_ = [
        ("A", vp), ("lda", C.c_int),
        ("W", vp),
        ("bias", vp),
        ("C", vp), ("ldc", C.c_int),
        ("M", C.c_int), ("N", C.c_int), ("K", C.c_int),
        ("act", C.c_int),
        ("rows_in", C.c_int), ("rows_out", C.c_int), ("row_off", C.c_int),
        ("addmat", vp),
        ("res", Rows),
        ("row_bias_scale", vp),
        ("preact", vp),
        ("aux", vp),
        ("split_hi", vp),
        ("split_lo", vp),
        ("in_T", C.c_int), ("in_off", C.c_int),
        ("addmat_stride", C.c_int64),
    ]


class BlockWeights(C.Structure):
    _fields_ = [(k, vp) for k in ("ln1_w", "ln1_b", "in_proj_w", "in_proj_b", "out_proj_w", "out_proj_b",
                                  "ln2_w", "ln2_b", "fc_w", "fc_b", "proj_w", "proj_b")]


class Instance(C.Structure):
    _fields_ = [
        ("L", C.c_int), ("n", C.c_int), ("M", C.c_int),
        ("nnz", C.c_int64),
        ("mip", vp), ("kpm", vp), ("pos_of_col", vp),
        ("csr_rowptr", vp), ("csr_col", vp), ("csr_val", vp),
        ("csc_colptr", vp), ("csc_row", vp), ("csc_val", vp),
        ("b", vp), ("c", vp),
        ("csr_val_int", vp), ("b_int", vp), ("c_int", vp),
    ]


class DdimCoeffs(C.Structure):
    _fields_ = [(k, C.c_float) for k in ("t", "sqrt_at", "s1m", "c_dir", "sqrt_aprev", "sigma")]


class DdpmCoeffs(C.Structure):
    _fields_ = [(k, C.c_float) for k in ("t", "c1", "c2", "std", "nonzero", "sqrt_recip", "sqrt_recipm1")]


i32, i64, f32, u64, sz = C.c_int, C.c_int64, C.c_float, C.c_uint64, C.c_size_t

# name -> (restype, argtypes); every symbol include/dip_b200.h declares
PROTOTYPES = {
    "dip_abi_version": (i32, []),
    "dip_last_error": (C.c_char_p, []),
    "dip_check_device": (i32, [i32, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32)]),
    "dip_coo_to_csr_csc_host": (i32, [i64, vp, vp, vp, i32, i32, vp, vp, vp, vp, vp, vp, C.POINTER(i64)]),
    "dip_ddim_step": (i32, [vp, vp, vp, vp, i64, vp, i64, vp, i32, i32, f32, f32, f32, f32, f32, f32, i32, vp]),
    "dip_ddpm_mean": (i32, [vp, vp, vp, i64, i32, i32, f32, f32, i32, f32, f32, vp]),
    "dip_ddpm_update": (i32, [vp, vp, vp, i64, vp, i32, i32, f32, f32, f32, vp]),
    "dip_randn": (i32, [vp, i64, u64, u64, vp]),
    "dip_gemm_nt": (i32, [C.POINTER(GemmArgs), vp]),
    "dip_layernorm_fwd": (i32, [vp, vp, vp, Rows, vp, vp, i64, vp]),
    "dip_layernorm_bwd": (i32, [vp, vp, Rows, vp, vp, vp, vp, i64, i32, vp]),
    "dip_attn_fwd": (i32, [vp, vp, vp, vp, vp, i32, vp, i64, i32, i32, vp]),
    "dip_attn_bwd": (i32, [vp, vp, vp, i32, vp, vp, vp, i32, vp, vp, vp, vp, i64, vp, i32, i32, vp]),
    "dip_split_bf16": (i32, [vp, vp, vp, i64, vp]),
    "dip_attn_fwd_tc": (i32, [vp, vp, vp, vp, vp, i64, i32, i32, i32, vp, i32, i32, vp]),
    "dip_attn_fwd_tc_state_bytes": (sz, [i32, i32]),
    "dip_attn_bwd_tc": (i32, [vp, vp, vp, i32, vp, vp, vp, vp, vp, vp, vp, vp, i64, vp, i32, i32, i32, i32, i32, vp, sz, vp]),
    "dip_attn_bwd_tc_ds_bytes": (sz, [i32, i32, i32, i32]),
    "dip_layernorm_split": (i32, [vp, vp, vp, vp, Rows, vp, vp, i64, i32, vp]),
    "dip_gemm_tc_planes": (i32, [C.POINTER(GemmArgs), vp, vp, vp, vp, vp]),
    "dip_gemm_tc": (i32, [C.POINTER(GemmArgs), vp, vp, C.POINTER(Rows), vp, vp, vp, vp, vp]),
    "dip_time_token": (i32, [vp, i32, i32, f32, vp, vp, vp]),
    "dip_block_chain_fwd_tc": (i32, [vp, vp, vp, vp, vp, vp, Rows, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, vp]),
    "dip_block_chain_bwd_tc": (i32, [vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, vp]),
    "dip_block_in_bwd_tc": (i32, [vp, vp, Rows, vp, vp, vp, vp, vp, vp, i32, i32, i32, vp]),
    "dip_decoder_head_fwd": (i32, [vp, vp, i32, i32, i32, vp, vp, vp, i64, vp]),
    "dip_decoder_head_bwd": (i32, [vp, vp, i32, i32, i32, vp, vp]),
    "dip_guidance": (i32, [vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, f32, i32, i32, i32, i32, i64, vp]),
    "dip_round_feasibility": (i32, [vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, i64, vp]),
    "dip_linear_smallk": (i32, [vp, vp, i64, i32, vp, vp, vp, vp, i32, vp]),
    "dip_conv_reduce": (i32, [vp, vp, vp, vp, vp, f32, f32, f32, vp, vp, vp, i32, vp]),
    "dip_spmm_csr": (i32, [vp, vp, vp, vp, vp, vp, i32, vp]),
    "dip_gather_pad": (i32, [vp, vp, vp, i32, i32, vp]),
    "dip_partial_select": (i32, [vp, vp, vp, i32, i32, i32, u64, u64, i32, i32, vp]),
    "dip_axpby": (i32, [vp, vp, vp, f32, f32, i64, vp]),
    "dip_engine_create": (i32, [C.POINTER(vp)]),
    "dip_engine_destroy": (None, [vp]),
    "dip_engine_set_denoiser": (i32, [vp, vp, vp, vp, vp, C.POINTER(BlockWeights), i32, vp]),
    "dip_engine_set_decoder": (i32, [vp, C.POINTER(BlockWeights), i32, vp, vp, vp]),
    "dip_engine_set_instance": (i32, [vp, C.POINTER(Instance), vp]),
    "dip_engine_set_instances": (i32, [vp, C.POINTER(Instance), i32, vp]),
    "dip_engine_set_wave_map": (i32, [vp, vp, i32, vp]),
    "dip_engine_n_max": (i32, [vp]),
    "dip_guidance_multi": (i32, [vp, vp, vp, vp, vp, vp, f32, i32, i32, i32, i32, vp]),
    "dip_round_feasibility_multi": (i32, [vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, vp]),
    "dip_engine_set_precision": (i32, [vp, i32]),
    "dip_engine_workspace_bytes": (sz, [vp, i32]),
    "dip_engine_reset_cache": (i32, [vp]),
    "dip_engine_denoise": (i32, [vp, vp, vp, i32, f32, vp, sz, vp]),
    "dip_engine_decode": (i32, [vp, vp, vp, vp, i32, vp, sz, vp]),
    "dip_engine_guidance_grad": (i32, [vp, vp, vp, vp, vp, vp, i32, f32, vp, sz, vp]),
    "dip_engine_guided_ddim_step": (i32, [vp, vp, vp, i32, C.POINTER(DdimCoeffs), f32, f32, i32, vp, vp, sz, vp]),
    "dip_engine_ddim_step": (i32, [vp, vp, vp, i32, C.POINTER(DdimCoeffs), i32, vp, vp, sz, vp]),
    "dip_engine_sample_guided_ddim": (i32, [vp, vp, i32, C.POINTER(DdimCoeffs), i32, f32, f32, i32, vp, vp, vp, vp, vp, sz, vp]),
    "dip_engine_final_decode": (i32, [vp, vp, i32, vp, vp, vp, vp, vp, vp, sz, vp]),
    "dip_engine_guided_ddpm_step": (i32, [vp, vp, i32, C.POINTER(DdpmCoeffs), f32, f32, i32, vp, vp, sz, vp]),
    "dip_engine_ddpm_step": (i32, [vp, vp, i32, C.POINTER(DdpmCoeffs), i32, vp, vp, sz, vp]),
    "dip_launch_count": (u64, []),
    "dip_prof_enable": (None, [i32]),
    "dip_prof_collect": (i32, [C.POINTER(C.c_double), C.POINTER(u64), i32]),
}

PROF_CLASSES = ["gemm", "layernorm", "attn_fwd", "attn_bwd_dkdv", "attn_bwd_dq", "guidance", "update", "head", "gnn", "other"]

_lib = None


def load():
    """Load libdip_b200.so (once).  Raises DipError when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise DipError("%s not found: run `python -c 'import __graft_entry__ as g; g.build()'` (nvcc, sm_100a). "
                       "There is no CPU or eager fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.dip_abi_version() != 2:
        raise DipError("libdip_b200.so ABI version mismatch")
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise DipError(load().dip_last_error().decode("utf-8", "replace"))


def ptr(t):
    """Device/host address of a torch tensor or numpy array (None -> NULL)."""
    if t is None:
        return None
    if hasattr(t, "data_ptr"):
        return t.data_ptr()
    return t.ctypes.data


def rows(batch, shared=None, T=0, split=0):
    return Rows(ptr(shared), ptr(batch), int(T), int(split))
