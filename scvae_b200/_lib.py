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
                                     c_f32, c_f32, c_f32, c_f32, c_ptr]),
    "scvae_step_advance": (c_int, [c_ptr, c_ptr]),
    "scvae_dp_reduce_adam": (c_int, [c_int, c_int, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_i64, c_ptr,
                                     c_f32, c_f32, c_f32, c_f32, c_f32, c_f32, c_ptr, c_int, c_ptr]),
    "scvae_group_offset_fwd": (c_int, [c_ptr, c_i64, c_ptr, c_i64, c_int, c_int, c_int, c_ptr, c_i64,
                                       c_ptr]),
    "scvae_group_offset_bwd": (c_int, [c_ptr, c_i64, c_int, c_int, c_int, c_ptr, c_i64, c_ptr, c_i64,
                                       c_int, c_ptr]),
    "scvae_softmax_fwd": (c_int, [c_ptr, c_i64, c_int, c_int, c_ptr, c_ptr, c_ptr]),
    "scvae_gmvae_latent_fwd": (c_int, [c_ptr, c_i64, c_ptr, c_int, c_int, c_int, c_int, c_ptr, c_ptr,
                                       c_i64, c_ptr, c_ptr, c_ptr]),
    "scvae_gmvae_latent_bwd": (c_int, [c_ptr, c_i64, c_ptr, c_int, c_int, c_int, c_int, c_ptr, c_ptr,
                                       c_i64, c_ptr, c_ptr, c_i64, c_ptr, c_ptr]),
    "scvae_gmvae_row_coefficients": (c_int, [c_ptr, c_int, c_int, c_int, c_f32, c_ptr, c_ptr, c_ptr]),
    "scvae_gmvae_bound": (c_int, [c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_int, c_int, c_int, c_f32,
                                  c_f32, c_int, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr]),
    "scvae_gmvae_z_mean": (c_int, [c_ptr, c_i64, c_ptr, c_int, c_int, c_int, c_ptr, c_ptr]),
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
