"""ctypes binding of ``libscvae_b200.so`` (the C ABI of ``include/scvae_b200.h``).

There is no CPU fallback: if the library cannot be built or loaded, importing the compute
path raises.  ``load()`` itself works without a GPU (the library links the CUDA runtime
statically and resolves the driver lazily), which is what the CPU-side symbol test uses.
"""

import ctypes
import os

from . import _build

c_int = ctypes.c_int
c_i64 = ctypes.c_int64
c_u64 = ctypes.c_uint64
c_f32 = ctypes.c_float
c_ptr = ctypes.c_void_p

MID_MAX_LAYERS = 4        # SCVAE_MID_MAX_LAYERS
_c_fp = ctypes.c_void_p   # float * / const float * / void * fields of the descriptor structs


class MidLayer(ctypes.Structure):
    """``scvae_mid_layer`` of include/scvae_b200.h, field by field."""
    _fields_ = [("w", _c_fp), ("dw", _c_fp), ("beta", _c_fp), ("dbeta", _c_fp),
                ("moving_mean", _c_fp), ("moving_var", _c_fp), ("mean", _c_fp), ("rstd", _c_fp),
                ("y", _c_fp), ("ldw", c_i64), ("ldy", c_i64),
                ("n_in", c_int), ("k_in", c_int), ("n_out", c_int), ("reserved", c_int)]


class MidDesc(ctypes.Structure):
    """``scvae_mid_desc`` of include/scvae_b200.h, field by field."""
    _fields_ = [("B", c_int), ("L", c_int), ("n_enc", c_int), ("n_dec", c_int),
                ("training", c_int), ("update_moving", c_int), ("deterministic", c_int),
                ("rows_per_cta", c_int),
                ("enc", MidLayer * MID_MAX_LAYERS), ("post", MidLayer),
                ("dec", MidLayer * MID_MAX_LAYERS),
                ("y1_parts", _c_fp), ("y1_ld", c_i64), ("y1_slice", c_i64), ("y1_nsplit", c_int),
                ("y1_alpha", c_f32),
                ("ph", _c_fp), ("ldph", c_i64),
                ("eps", _c_fp), ("generate_eps", c_int), ("reserved0", c_int), ("seed", c_u64),
                ("offset", c_u64), ("offset_dev", _c_fp),
                ("z", _c_fp), ("ldz", c_i64),
                ("kl_row", _c_fp), ("kl_elem", _c_fp),
                ("batch_index", _c_fp), ("count_sum", _c_fp), ("n_batches", c_int),
                ("reserved1", c_int),
                ("d16", _c_fp), ("ldd16", c_i64),
                ("h_last", _c_fp), ("ldh_last", c_i64),
                ("dd_parts", _c_fp), ("dd_ld", c_i64), ("dd_slice", c_i64), ("dd_nsplit", c_int),
                ("logp_nsplit", c_int),
                ("logp_parts", _c_fp), ("logp_slice", c_i64), ("row_const", _c_fp),
                ("logp", _c_fp), ("bound", _c_fp),
                ("dy1_16", _c_fp), ("dy1_16_lo", _c_fp), ("lddy1", c_i64), ("dy1", _c_fp),
                ("lddy1_f32", c_i64),
                ("go_scalar", c_f32), ("dy1_scale", c_f32), ("kl_weight", c_f32),
                ("reserved2", c_int), ("scalars", _c_fp),
                ("workspace", _c_fp), ("workspace_floats", c_i64), ("barrier", _c_fp),
                ("error", _c_fp), ("timeline", _c_fp)]


class Shadow(ctypes.Structure):
    """``scvae_shadow`` of include/scvae_b200.h."""
    _fields_ = [("lo", c_i64), ("hi", c_i64), ("src_ld", c_i64), ("dst_ld", c_i64),
                ("src_block_rows", c_i64), ("dst_block_rows", c_i64), ("hi16", _c_fp),
                ("lo16", _c_fp), ("cols", c_int), ("reserved", c_int)]


# name -> (restype, argtypes); mirrors include/scvae_b200.h declaration by declaration.
SIGNATURES = {
    "scvae_abi_version": (c_int, []),
    "scvae_last_error": (ctypes.c_char_p, []),
    "scvae_num_heads": (c_int, [c_int]),
    "scvae_launch_count": (ctypes.c_longlong, []),
    "scvae_csr_densify": (c_int, [c_ptr, c_ptr, c_ptr, c_ptr, c_int, c_int, c_ptr, c_i64, c_ptr,
                                  c_int, c_ptr, c_i64, c_ptr, c_i64, c_ptr]),
    "scvae_csr_densify_u16": (c_int, [c_ptr, c_ptr, c_ptr, c_ptr, c_int, c_int, c_ptr, c_i64, c_ptr,
                                      c_int, c_ptr, c_i64, c_ptr, c_i64, c_ptr]),
    "scvae_packed_rows_offset": (c_i64, [c_int]),
    "scvae_csr_densify_packed": (c_int, [c_ptr, c_int, c_int, c_ptr, c_ptr, c_i64, c_ptr, c_i64, c_ptr]),
    "scvae_packed_pull": (c_int, [c_ptr, c_ptr, c_ptr, c_ptr, c_int, c_ptr, c_i64, c_ptr]),
    "scvae_packed_copy_batch": (c_int, [c_ptr, c_ptr, c_ptr, c_ptr, c_int, c_i64, c_ptr, c_ptr, c_i64, c_int,
                                        c_ptr, c_ptr]),
    "scvae_pack_row_slab": (c_int, [c_ptr, c_ptr, c_ptr, c_ptr, c_int, c_i64, c_ptr, c_i64, c_int, c_ptr]),
    "scvae_f32_to_u16": (c_int, [c_ptr, c_i64, c_i64, c_int, c_ptr, c_i64, c_ptr]),
    "scvae_csr_row_constants": (c_int, [c_ptr, c_ptr, c_int, c_i64, c_ptr, c_ptr]),
    "scvae_gather_f32": (c_int, [c_ptr, c_ptr, c_int, c_ptr, c_ptr]),
    "scvae_heads_fused_workspace_floats": (c_i64, [c_int, c_int]),
    "scvae_heads_fused_fwd": (c_int, [c_int, c_ptr, c_ptr, c_i64, c_ptr, c_i64, c_int, c_int, c_int, c_int,
                                      c_ptr, c_ptr, c_ptr, c_ptr]),
    "scvae_heads_fused_bwd": (c_int, [c_int, c_ptr, c_ptr, c_i64, c_ptr, c_i64, c_int, c_int, c_int, c_int,
                                      c_ptr, c_ptr, c_f32, c_f32, c_ptr, c_ptr, c_i64, c_int, c_ptr,
                                      c_ptr, c_ptr]),
    "scvae_gemm_f32": (c_int, [c_int, c_int, c_int, c_int, c_ptr, c_i64, c_ptr, c_i64, c_ptr,
                               c_i64, c_int, c_ptr]),
    "scvae_gemm_tf32": (c_int, [c_int, c_int, c_int, c_int, c_ptr, c_i64, c_ptr, c_i64, c_ptr,
                                c_i64, c_int, c_ptr, c_i64, c_ptr]),
    "scvae_gemm_tf32_workspace_bytes": (c_i64, [c_int, c_int, c_int, c_int]),
    "scvae_gemm_f16": (c_int, [c_int, c_int, c_int, c_int, c_ptr, c_i64, c_ptr, c_i64, c_ptr, c_i64,
                               c_int, c_f32, c_ptr, c_i64, c_ptr]),
    "scvae_gemm_f16_workspace_bytes": (c_i64, [c_int, c_int, c_int, c_int]),
    "scvae_gemm_f16_split": (c_int, [c_int, c_int, c_int, c_int, c_ptr, c_i64, c_ptr, c_i64, c_ptr, c_i64,
                                     c_int, c_ptr, c_i64, c_int, c_f32, c_ptr, c_i64, c_ptr]),
    "scvae_f32_to_f16_split": (c_int, [c_ptr, c_i64, c_i64, c_int, c_ptr, c_ptr, c_i64, c_f32, c_ptr]),
    "scvae_gemm_sm_limit": (c_int, [c_int]),
    "scvae_f32_to_f16": (c_int, [c_ptr, c_i64, c_i64, c_int, c_ptr, c_i64, c_f32, c_ptr]),
    "scvae_bn_scratch_floats": (c_i64, [c_int, c_int, c_int]),
    "scvae_bn_act_fwd": (c_int, [c_ptr, c_i64, c_int, c_int, c_int, c_ptr, c_ptr, c_ptr, c_int,
                                 c_int, c_int, c_ptr, c_i64, c_ptr, c_ptr, c_ptr, c_ptr]),
    "scvae_bn_act_bwd": (c_int, [c_ptr, c_i64, c_ptr, c_i64, c_ptr, c_i64, c_int, c_int, c_int,
                                 c_ptr, c_ptr, c_int, c_ptr, c_i64, c_ptr, c_int, c_ptr, c_ptr]),
    "scvae_act_fwd": (c_int, [c_ptr, c_i64, c_int, c_int, c_int, c_ptr, c_i64, c_ptr]),
    "scvae_act_bwd": (c_int, [c_ptr, c_i64, c_ptr, c_i64, c_int, c_int, c_int, c_ptr, c_i64,
                              c_ptr]),
    "scvae_gaussian_latent_fwd": (c_int, [c_ptr, c_i64, c_int, c_int, c_int, c_ptr, c_int, c_int,
                                          c_ptr, c_i64, c_ptr, c_ptr, c_ptr]),
    "scvae_gaussian_latent_bwd": (c_int, [c_ptr, c_i64, c_int, c_int, c_int, c_ptr, c_int, c_ptr,
                                          c_i64, c_f32, c_ptr, c_i64, c_ptr]),
    "scvae_gaussian_sampled_kl": (c_int, [c_ptr, c_i64, c_int, c_int, c_int, c_ptr, c_int, c_int,
                                          c_ptr, c_ptr, c_ptr]),
    "scvae_gaussian_sampled_kl_bwd": (c_int, [c_ptr, c_i64, c_int, c_int, c_int, c_ptr, c_int,
                                              c_ptr, c_i64, c_ptr, c_f32, c_f32, c_ptr, c_i64,
                                              c_ptr]),
    "scvae_vae_bound_rows": (c_int, [c_ptr, c_ptr, c_int, c_int, c_int, c_f32, c_ptr, c_ptr,
                                     c_ptr]),
    "scvae_dropout_fwd": (c_int, [c_ptr, c_i64, c_int, c_int, c_int, c_ptr, c_f32, c_f32, c_ptr,
                                  c_i64, c_int, c_ptr]),
    "scvae_dropout_bwd": (c_int, [c_ptr, c_i64, c_int, c_int, c_int, c_ptr, c_f32, c_f32, c_ptr,
                                  c_i64, c_int, c_ptr]),
    "scvae_likelihood_fwd": (c_int, [c_int, c_ptr, c_i64, c_int, c_ptr, c_i64, c_i64, c_int,
                                     c_int, c_ptr, c_ptr, c_ptr]),
    "scvae_likelihood_bwd": (c_int, [c_int, c_ptr, c_i64, c_int, c_ptr, c_i64, c_i64, c_int,
                                     c_int, c_ptr, c_ptr, c_f32, c_ptr, c_i64, c_i64, c_ptr,
                                     c_ptr]),
    "scvae_likelihood_moments": (c_int, [c_int, c_ptr, c_i64, c_i64, c_int, c_int, c_int, c_int,
                                         c_ptr, c_i64, c_ptr, c_ptr, c_ptr, c_i64, c_ptr]),
    "scvae_continuous_num_heads": (c_int, [c_int]),
    "scvae_continuous_likelihood": (c_int, [c_int, c_ptr, c_i64, c_int, c_ptr, c_i64, c_i64, c_int,
                                            c_int, c_ptr, c_f32, c_ptr, c_i64, c_i64, c_ptr, c_ptr]),
    "scvae_continuous_moments": (c_int, [c_int, c_ptr, c_i64, c_i64, c_int, c_int, c_int, c_int,
                                         c_ptr, c_i64, c_ptr, c_ptr, c_ptr, c_i64, c_ptr]),
    "scvae_decoder_features": (c_int, [c_ptr, c_i64, c_int, c_int, c_int, c_ptr, c_int, c_ptr, c_ptr]),
    "scvae_piecewise_likelihood": (c_int, [c_int, c_int, c_ptr, c_i64, c_int, c_ptr, c_i64, c_i64, c_int,
                                           c_int, c_ptr, c_f32, c_ptr, c_i64, c_i64, c_ptr, c_ptr]),
    "scvae_piecewise_moments": (c_int, [c_int, c_int, c_ptr, c_i64, c_i64, c_int, c_int, c_int, c_int,
                                        c_ptr, c_i64, c_ptr, c_ptr, c_ptr, c_i64, c_ptr]),
    "scvae_constrained_poisson": (c_int, [c_ptr, c_i64, c_int, c_ptr, c_i64, c_int, c_int, c_ptr, c_ptr,
                                          c_ptr, c_f32, c_ptr, c_i64, c_ptr, c_ptr, c_ptr]),
    "scvae_constrained_poisson_mixture_moments": (c_int, [c_ptr, c_i64, c_ptr, c_ptr, c_int, c_int,
                                                          c_int, c_int, c_ptr, c_i64, c_ptr, c_ptr,
                                                          c_ptr, c_i64, c_ptr]),
    "scvae_constrained_poisson_moments": (c_int, [c_ptr, c_i64, c_ptr, c_ptr, c_int, c_int, c_int, c_ptr,
                                                  c_ptr, c_ptr, c_i64, c_ptr]),
    "scvae_vae_bound": (c_int, [c_ptr, c_ptr, c_int, c_int, c_int, c_f32, c_ptr, c_ptr, c_ptr]),
    "scvae_adam_clip_step": (c_int, [c_ptr, c_ptr, c_ptr, c_ptr, c_i64, c_ptr, c_f32, c_f32,
                                     c_f32, c_f32, c_f32, c_f32, c_ptr, ctypes.POINTER(Shadow), c_int,
                                     c_ptr, c_int, c_ptr]),
    "scvae_adam_clip_ctas": (c_int, [c_i64]),
    "scvae_step_advance": (c_int, [c_ptr, c_ptr]),
    "scvae_dp_reduce_adam": (c_int, [c_int, c_int, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_i64, c_ptr,
                                     c_f32, c_f32, c_f32, c_f32, c_f32, c_f32, c_ptr, c_ptr, c_int,
                                     c_ptr]),
    "scvae_group_offset_fwd": (c_int, [c_ptr, c_i64, c_ptr, c_i64, c_int, c_int, c_int, c_ptr, c_i64,
                                       c_ptr]),
    "scvae_group_offset_bwd": (c_int, [c_ptr, c_i64, c_int, c_int, c_int, c_ptr, c_i64, c_ptr, c_i64,
                                       c_int, c_ptr]),
    "scvae_softmax_fwd": (c_int, [c_ptr, c_i64, c_int, c_int, c_ptr, c_ptr, c_ptr]),
    "scvae_gmvae_latent_fwd": (c_int, [c_ptr, c_i64, c_ptr, c_int, c_int, c_int, c_int, c_ptr, c_ptr,
                                       c_i64, c_ptr, c_ptr, c_ptr]),
    "scvae_gmvae_full_prior": (c_int, [c_ptr, c_i64, c_int, c_int, c_ptr, c_ptr]),
    "scvae_gmvae_latent_full_fwd": (c_int, [c_ptr, c_i64, c_ptr, c_i64, c_ptr, c_int, c_int, c_int, c_int, c_ptr,
                                            c_ptr, c_i64, c_ptr, c_ptr, c_ptr]),
    "scvae_gmvae_latent_full_bwd": (c_int, [c_ptr, c_i64, c_ptr, c_i64, c_ptr, c_int, c_int, c_int, c_int, c_ptr,
                                            c_ptr, c_i64, c_ptr, c_ptr, c_ptr, c_ptr, c_i64, c_ptr, c_i64, c_ptr]),
    "scvae_gmvae_full_covariance_mean": (c_int, [c_ptr, c_i64, c_int, c_int, c_int, c_ptr, c_ptr]),
    "scvae_gmvae_latent_bwd": (c_int, [c_ptr, c_i64, c_ptr, c_int, c_int, c_int, c_int, c_ptr, c_ptr,
                                       c_i64, c_ptr, c_ptr, c_i64, c_ptr, c_ptr]),
    "scvae_gmvae_row_coefficients": (c_int, [c_ptr, c_int, c_int, c_int, c_f32, c_ptr, c_ptr, c_ptr]),
    "scvae_gmvae_bound": (c_int, [c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_int, c_int, c_int, c_f32,
                                  c_f32, c_int, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr]),
    "scvae_gmvae_z_mean": (c_int, [c_ptr, c_i64, c_ptr, c_int, c_int, c_int, c_ptr, c_ptr]),
    "scvae_vae_mid_workspace_floats": (c_i64, [ctypes.POINTER(MidDesc)]),
    "scvae_vae_mid_fwd": (c_int, [ctypes.POINTER(MidDesc), c_ptr]),
    "scvae_vae_mid_bwd": (c_int, [ctypes.POINTER(MidDesc), c_ptr]),
    "scvae_col_mean": (c_int, [c_ptr, c_i64, c_int, c_int, c_ptr, c_ptr]),
    "scvae_fill_normal": (c_int, [c_ptr, c_i64, c_u64, c_u64, c_ptr, c_ptr]),
}

_LIB = None


class ScvaeNativeError(RuntimeError):
    pass


def load(rebuild_if_stale=True):
    """Load (building first if needed) the native library; raises if impossible."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = _build.LIB_PATH
    if rebuild_if_stale and _build.find_nvcc() is not None:
        path = _build.build()
    if not os.path.exists(path):
        raise ScvaeNativeError(
            "libscvae_b200.so is missing and nvcc is unavailable: the scVAE hot path has no "
            "CPU fallback. Run `python -m scvae_b200._build` on a machine with CUDA 12.9.")
    lib = ctypes.CDLL(path)
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = restype
        fn.argtypes = argtypes
    if lib.scvae_abi_version() != 1:
        raise ScvaeNativeError("libscvae_b200.so ABI version mismatch")
    _LIB = lib
    return lib


def check(status, what=""):
    if status != 0:
        msg = load().scvae_last_error().decode("utf-8", "replace")
        raise ScvaeNativeError("{} failed ({}): {}".format(what or "native call", status, msg))
