"""Thin tensor-facing wrappers over the C ABI (``include/scvae_b200.h``).

PyTorch is used only for device memory and streams: every function here takes CUDA tensors,
passes raw device pointers + the current stream to ``libscvae_b200.so`` and returns nothing
(outputs are caller-allocated).  No function falls back to torch math.
"""

import torch

from . import _lib

LIKELIHOOD_KINDS = {
    "poisson": 0,
    "negative binomial": 1,
    "zero-inflated poisson": 2,
    "zero-inflated negative binomial": 3,
    "constrained poisson": 4,      # row kernel of its own (softmax over genes), not in the family
    # continuous / binary reconstruction distributions (csrc/continuous.cu)
    "gaussian": 8,
    "softplus gaussian": 9,
    "modified gaussian": 9,        # DU:306: an alias
    "log-normal": 10,
    "gamma": 11,
    "bernoulli": 12,
    "lomax": 13,
    "exponentially_modified_gaussian": 14,
}
CONSTRAINED_POISSON = 4
CONTINUOUS_KINDS = frozenset(range(8, 15))
LIKELIHOOD_HEADS = {
    "constrained poisson": ["lambda"],
    "poisson": ["log_lambda"],
    "negative binomial": ["p", "log_r"],
    "zero-inflated poisson": ["pi", "log_lambda"],
    "zero-inflated negative binomial": ["pi", "p", "log_r"],
    "gaussian": ["mu", "log_sigma"],
    "softplus gaussian": ["mean", "softplus_scale"],
    "modified gaussian": ["mean", "softplus_scale"],
    "log-normal": ["mean", "variance"],
    "gamma": ["concentration", "rate"],
    "bernoulli": ["logits"],
    "lomax": ["log_concentration", "log_scale"],
    "exponentially_modified_gaussian": ["location", "scale", "rate"],
}
GEMM_NT, GEMM_NN, GEMM_TN = 0, 1, 2


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _p(t):
    if t is None:
        return None
    if not t.is_cuda:
        raise _lib.ScvaeNativeError("scvae_b200 kernels need CUDA tensors (no CPU fallback)")
    return t.data_ptr()


def _f32(*ts):
    for t in ts:
        if t is not None and t.dtype != torch.float32:
            raise TypeError("expected float32 tensor, got {}".format(t.dtype))


def _ld(t):
    """Leading dimension (elements) of a 2-D row-major view."""
    if t.dim() != 2 or t.stride(1) != 1:
        raise ValueError("expected a 2-D row-major (unit inner stride) tensor")
    return t.stride(0)


def csr_densify(indptr, indices, values, rows, G, x, row_const=None, rebase=False, t16=None,
                x16=None):
    """CSR rows -> dense minibatch; any of x (fp32), t16 (uint16), x16 (fp16) may be None."""
    B = next(t for t in (x, t16, x16) if t is not None).shape[0]
    lib = _lib.load()
    compact = indices.dtype == torch.int16
    if compact != (values.dtype == torch.int16):
        raise TypeError("compact CSR needs both indices and values as uint16 (int16 storage)")
    fn = lib.scvae_csr_densify_u16 if compact else lib.scvae_csr_densify
    _lib.check(fn(_p(indptr), _p(indices), _p(values), _p(rows), B, G, _p(x),
                                     _ld(x) if x is not None else 0, _p(row_const), int(rebase),
                                     _p(t16), _ld(t16) if t16 is not None else 0, _p(x16),
                                     _ld(x16) if x16 is not None else 0, _stream()),
               "csr_densify")


def csr_densify_packed(slab, B, G, row_const=None, t16=None, x16=None):
    """Minibatch from one packed row slab (layout: include/scvae_b200.h, hotloop.PackedStream)."""
    lib = _lib.load()
    _lib.check(lib.scvae_csr_densify_packed(_p(slab), B, G, _p(row_const), _p(t16),
                                            _ld(t16) if t16 is not None else 0, _p(x16),
                                            _ld(x16) if x16 is not None else 0, _stream()),
               "csr_densify_packed")


def packed_pull(store_pinned, row_off, row_const_all, order, slab):
    """The GPU assembles the packed slab of the rows ``order`` (device int64) itself, reading the
    row strings from pinned host memory (``store_pinned``: a pinned uint8 tensor)."""
    lib = _lib.load()
    _lib.check(lib.scvae_packed_pull(store_pinned.data_ptr(), _p(row_off), _p(row_const_all), _p(order),
                                     int(order.numel()), _p(slab), int(slab.numel()), _stream()),
               "packed_pull")


def packed_copy_batch(store, row_off, row_const_all, order, header, slab):
    """One batched copy (copy engine) of the rows ``order`` from the pinned ``store`` (numpy view of
    pinned memory) into the device ``slab``; ``header``: numpy view of a pinned staging area.
    Returns the slab's bytes, or None when the runtime has no batched copies."""
    import ctypes
    lib = _lib.load()
    out = ctypes.c_int64(0)
    status = lib.scvae_packed_copy_batch(store.ctypes.data, row_off.ctypes.data, row_const_all.ctypes.data,
                                         order.ctypes.data, int(order.size), int(row_off.size - 1),
                                         header.ctypes.data, _p(slab), int(slab.numel()),
                                         slab.device.index or 0, _stream(), ctypes.byref(out))
    if status == 2:
        return None
    _lib.check(status, "packed_copy_batch")
    return int(out.value)


def packed_rows_offset(B):
    return int(_lib.load().scvae_packed_rows_offset(int(B)))


def pack_row_slab(store, row_off, row_const_all, order, dst, threads=4):
    """HOST: assemble the packed slab of the rows ``order`` in ``dst`` (numpy uint8, e.g. a view of
    pinned memory) from the per-row strings ``store`` / ``row_off``; returns the slab's bytes.
    The call releases the GIL (ctypes), so a feeder thread packs ahead of the GPU."""
    import ctypes
    lib = _lib.load()
    out = ctypes.c_int64(0)
    _lib.check(lib.scvae_pack_row_slab(store.ctypes.data, row_off.ctypes.data, row_const_all.ctypes.data,
                                       order.ctypes.data, int(order.size), int(row_off.size - 1),
                                       dst.ctypes.data, int(dst.size), int(threads), ctypes.byref(out)),
               "pack_row_slab")
    return int(out.value)


def csr_row_constants(indptr, values, out):
    """out[r] = sum_g lgamma(1 + x[r, g]) for every CSR row (absolute indptr)."""
    lib = _lib.load()
    _lib.check(lib.scvae_csr_row_constants(_p(indptr), _p(values), int(values.dtype == torch.int16),
                                           out.numel(), _p(out), _stream()), "csr_row_constants")


def gather_f32(src, rows, dst):
    lib = _lib.load()
    _lib.check(lib.scvae_gather_f32(_p(src), _p(rows), dst.numel(), _p(dst), _stream()), "gather_f32")


def f32_to_u16(x, G, t16):
    lib = _lib.load()
    _lib.check(lib.scvae_f32_to_u16(_p(x), _ld(x), x.shape[0], G, _p(t16), _ld(t16), _stream()),
               "f32_to_u16")


def heads_fused_workspace_floats(M, G):
    return int(_lib.load().scvae_heads_fused_workspace_floats(M, G))


def heads_fused_fwd(kind, d16, w16, head_stride, t16, M, G, logp, workspace, row_const=None):
    """Forward-only fused heads GEMM + likelihood (evaluation passes)."""
    lib = _lib.load()
    _lib.check(lib.scvae_heads_fused_fwd(kind, _p(d16), _p(w16), head_stride, _p(t16), _ld(t16),
                                         int(t16.dtype == torch.float16), t16.shape[0], M, G,
                                         _p(row_const), _p(logp), _p(workspace), _stream()),
               "heads_fused_fwd")


def heads_fused_bwd(kind, d16, w16, head_stride, t16, M, G, da16, dd, dd_cols, logp, workspace,
                    row_const=None, go=None, go_scalar=1.0, scale=1.0):
    """Fused heads GEMM + likelihood forward/backward + decoder-gradient GEMM.  ``t16`` is a
    uint16 (torch.int16 storage) or a torch.float16 target matrix."""
    lib = _lib.load()
    _lib.check(lib.scvae_heads_fused_bwd(kind, _p(d16), _p(w16), head_stride, _p(t16), _ld(t16),
                                         int(t16.dtype == torch.float16), t16.shape[0], M, G, _p(row_const), _p(go),
                                         float(go_scalar), float(scale), _p(da16), _p(dd), _ld(dd),
                                         dd_cols, _p(logp), _p(workspace), _stream()),
               "heads_fused_bwd")


def continuous_likelihood(kind, t, a, head_stride, M, G, logp=None, go=None, go_scalar=1.0, da=None):
    """log p (and, with ``da``, its gradient) of a continuous / binary reconstruction distribution."""
    _f32(t, a)
    lib = _lib.load()
    _lib.check(lib.scvae_continuous_likelihood(kind, _p(t), _ld(t), t.shape[0], _p(a), _ld(a),
                                               head_stride, M, G, _p(go), float(go_scalar), _p(da),
                                               _ld(da) if da is not None else 0, head_stride,
                                               _p(logp), _stream()), "continuous_likelihood")


def _ldo(*outs):
    """Row pitch shared by the moment outputs that are requested (None = skipped)."""
    given = [o for o in outs if o is not None]
    if not given or any(_ld(o) != _ld(given[0]) for o in given):
        raise ValueError("moment outputs: at least one, all with the same row pitch")
    return _ld(given[0])


def continuous_moments(kind, a, head_stride, B, G, RS, K_, y, p_x_mean, p_x_stddev, stddev_of_mean):
    lib = _lib.load()
    _lib.check(lib.scvae_continuous_moments(kind, _p(a), _ld(a), head_stride, B, G, RS, K_, _p(y),
                                            _ld(y) if y is not None else 0, _p(p_x_mean),
                                            _p(p_x_stddev), _p(stddev_of_mean), _ldo(p_x_mean, p_x_stddev, stddev_of_mean),
                                            _stream()), "continuous_moments")


def gemm(layout, M, N, K, A, B, C, accumulate=False, tensor_cores=True, workspace=None):
    """C[M,N] (+)= op(A) op(B).  A, B, C are 2-D row-major views (only ld and ptr are used)."""
    _f32(A, B, C)
    lib = _lib.load()
    if tensor_cores:
        ws_ptr = _p(workspace)
        ws_bytes = workspace.numel() * workspace.element_size() if workspace is not None else 0
        _lib.check(lib.scvae_gemm_tf32(layout, M, N, K, _p(A), _ld(A), _p(B), _ld(B), _p(C),
                                       _ld(C), int(accumulate), ws_ptr, ws_bytes, _stream()),
                   "gemm_tf32")
    else:
        _lib.check(lib.scvae_gemm_f32(layout, M, N, K, _p(A), _ld(A), _p(B), _ld(B), _p(C),
                                      _ld(C), int(accumulate), _stream()), "gemm_f32")


def gemm_f16_split(layout, M, N, K, A, B, X, which, C, accumulate=False, alpha=1.0, workspace=None):
    """fp16 product with operand ``which`` (1 = A, 2 = B) completed by its rounding remainder X."""
    lib = _lib.load()
    ws_bytes = workspace.numel() * workspace.element_size() if workspace is not None else 0
    _lib.check(lib.scvae_gemm_f16_split(layout, M, N, K, _p(A), _ld(A), _p(B), _ld(B), _p(X), _ld(X),
                                        int(which), _p(C), _ld(C), int(accumulate), float(alpha),
                                        _p(workspace), ws_bytes, _stream()), "gemm_f16_split")


def f32_to_f16_split(src, cols, hi, lo, scale=1.0):
    lib = _lib.load()
    _lib.check(lib.scvae_f32_to_f16_split(_p(src), _ld(src), src.shape[0], cols, _p(hi), _p(lo),
                                          _ld(hi), float(scale), _stream()), "f32_to_f16_split")


def gemm_f16(layout, M, N, K, A, B, C, accumulate=False, alpha=1.0, workspace=None):
    """fp16 operands (torch.float16 2-D row-major views), fp32 output."""
    lib = _lib.load()
    ws_bytes = workspace.numel() * workspace.element_size() if workspace is not None else 0
    _lib.check(lib.scvae_gemm_f16(layout, M, N, K, _p(A), _ld(A), _p(B), _ld(B), _p(C), _ld(C),
                                  int(accumulate), float(alpha), _p(workspace), ws_bytes,
                                  _stream()), "gemm_f16")


def gemm_sm_limit(max_ctas):
    """Bound the persistent CTAs of subsequent tensor-core GEMMs (0 = all SMs); returns the old bound."""
    return int(_lib.load().scvae_gemm_sm_limit(int(max_ctas)))


def gemm_f16_workspace_bytes(layout, M, N, K):
    return int(_lib.load().scvae_gemm_f16_workspace_bytes(layout, M, N, K))


def f32_to_f16(src, cols, dst, scale=1.0):
    lib = _lib.load()
    _lib.check(lib.scvae_f32_to_f16(_p(src), _ld(src), src.shape[0], cols, _p(dst), _ld(dst),
                                    float(scale), _stream()), "f32_to_f16")


def gemm_workspace_bytes(layout, M, N, K):
    return int(_lib.load().scvae_gemm_tf32_workspace_bytes(layout, M, N, K))


def bn_scratch_floats(M, H, groups=1):
    return int(_lib.load().scvae_bn_scratch_floats(M, H, groups))


def bn_act_fwd(y, H, beta, moving_mean, moving_var, out, save_mean, save_rstd, scratch,
               training=True, update_moving=True, relu=True, groups=1):
    lib = _lib.load()
    _lib.check(lib.scvae_bn_act_fwd(_p(y), _ld(y), y.shape[0], H, groups, _p(beta),
                                    _p(moving_mean), _p(moving_var), int(training),
                                    int(update_moving), int(relu), _p(out), _ld(out),
                                    _p(save_mean), _p(save_rstd), _p(scratch), _stream()),
               "bn_act_fwd")


def bn_act_bwd(dout, y, out, H, save_mean, save_rstd, dy, dbeta, scratch, relu=True, groups=1,
               accumulate_dbeta=False):
    lib = _lib.load()
    _lib.check(lib.scvae_bn_act_bwd(_p(dout), _ld(dout), _p(y), _ld(y), _p(out), _ld(out),
                                    y.shape[0], H, groups, _p(save_mean), _p(save_rstd),
                                    int(relu), _p(dy), _ld(dy), _p(dbeta), int(accumulate_dbeta),
                                    _p(scratch), _stream()), "bn_act_bwd")


def act_fwd(y, H, out, relu=True):
    lib = _lib.load()
    _lib.check(lib.scvae_act_fwd(_p(y), _ld(y), y.shape[0], H, int(relu), _p(out), _ld(out),
                                 _stream()), "act_fwd")


def act_bwd(dout, out, H, dy, relu=True):
    lib = _lib.load()
    _lib.check(lib.scvae_act_bwd(_p(dout), _ld(dout), _p(out), _ld(out), out.shape[0], H,
                                 int(relu), _p(dy), _ld(dy), _stream()), "act_bwd")


def gaussian_latent_fwd(ph, B, L, RS, eps, z, kl_row, kl_elem=None, unit_variance=False,
                        deterministic=False):
    lib = _lib.load()
    _lib.check(lib.scvae_gaussian_latent_fwd(_p(ph), _ld(ph), B, L, RS, _p(eps),
                                             int(unit_variance), int(deterministic), _p(z),
                                             _ld(z), _p(kl_row), _p(kl_elem), _stream()),
               "gaussian_latent_fwd")


def decoder_features(z, M, B, col0, batch_index=None, n_batches=0, count_sum=None):
    """One-hot batch index / count-sum columns of the decoder input (behind the ones column)."""
    lib = _lib.load()
    _lib.check(lib.scvae_decoder_features(_p(z), _ld(z), M, B, col0, _p(batch_index), n_batches,
                                          _p(count_sum), _stream()), "decoder_features")


def gaussian_latent_bwd(ph, B, L, RS, eps, dz, kl_coef, dph, unit_variance=False):
    lib = _lib.load()
    _lib.check(lib.scvae_gaussian_latent_bwd(_p(ph), _ld(ph), B, L, RS, _p(eps),
                                             int(unit_variance), _p(dz), _ld(dz), float(kl_coef),
                                             _p(dph), _ld(dph), _stream()),
               "gaussian_latent_bwd")


def dropout_fwd(x, rows, n, skip_col, noise, threshold, keep, out, width):
    """out = x * [noise < threshold] / keep on the logical columns [0, n) (skipping the physical
    column ``skip_col``), other stored columns copied (MU:45-50)."""
    lib = _lib.load()
    _lib.check(lib.scvae_dropout_fwd(_p(x), _ld(x), rows, n, skip_col, _p(noise), float(threshold),
                                     float(keep), _p(out), _ld(out), width, _stream()),
               "dropout_fwd")


def dropout_bwd(dx, rows, n, skip_col, noise, threshold, keep, dsrc=None, accumulate=False):
    lib = _lib.load()
    _lib.check(lib.scvae_dropout_bwd(_p(dx), _ld(dx), rows, n, skip_col, _p(noise),
                                     float(threshold), float(keep), _p(dsrc),
                                     _ld(dsrc) if dsrc is not None else 0, int(accumulate),
                                     _stream()), "dropout_bwd")


def gaussian_sampled_kl(ph, B, L, RS, eps, kl_rows, kl_elem=None, unit_variance=False,
                        deterministic=False):
    """Sampled KL term per (sample, cell) row (VAE:2628-2640)."""
    lib = _lib.load()
    _lib.check(lib.scvae_gaussian_sampled_kl(_p(ph), _ld(ph), B, L, RS, _p(eps),
                                             int(unit_variance), int(deterministic), _p(kl_rows),
                                             _p(kl_elem), _stream()), "gaussian_sampled_kl")


def gaussian_sampled_kl_bwd(ph, B, L, RS, eps, dz, go, weight, coef_scalar, dph,
                            unit_variance=False):
    lib = _lib.load()
    _lib.check(lib.scvae_gaussian_sampled_kl_bwd(_p(ph), _ld(ph), B, L, RS, _p(eps),
                                                 int(unit_variance), _p(dz), _ld(dz), _p(go),
                                                 float(weight), float(coef_scalar), _p(dph),
                                                 _ld(dph), _stream()), "gaussian_sampled_kl_bwd")


def vae_bound_rows(logp, kl_rows, R, S, B, weight, out, go=None):
    lib = _lib.load()
    _lib.check(lib.scvae_vae_bound_rows(_p(logp), _p(kl_rows), R, S, B, float(weight), _p(out),
                                        _p(go), _stream()), "vae_bound_rows")


def piecewise_likelihood(kind, k_max, t, a, head_stride, M, G, logp=None, go=None, go_scalar=1.0,
                         da=None):
    """Categorised count likelihood (`-k`): P heads of ``kind`` + k_max + 1 class-logit heads."""
    _f32(t, a)
    lib = _lib.load()
    _lib.check(lib.scvae_piecewise_likelihood(kind, k_max, _p(t), _ld(t), t.shape[0], _p(a), _ld(a),
                                              head_stride, M, G, _p(go), float(go_scalar), _p(da),
                                              _ld(da) if da is not None else 0, head_stride,
                                              _p(logp), _stream()), "piecewise_likelihood")


def piecewise_moments(kind, k_max, a, head_stride, B, G, RS, p_x_mean, p_x_stddev, stddev_of_mean,
                      K_=1, y=None):
    lib = _lib.load()
    _lib.check(lib.scvae_piecewise_moments(kind, k_max, _p(a), _ld(a), head_stride, B, G, RS, K_,
                                           _p(y), _ld(y) if y is not None else 0, _p(p_x_mean),
                                           _p(p_x_stddev), _p(stddev_of_mean), _ldo(p_x_mean, p_x_stddev, stddev_of_mean),
                                           _stream()), "piecewise_moments")


def constrained_poisson(t, a, M, G, count_sum, logp=None, row_const=None, go=None, go_scalar=1.0,
                        da=None, lse=None):
    """Constrained Poisson log p (and gradient when ``da`` is given); rows of t tile over M."""
    _f32(t, a)
    lib = _lib.load()
    _lib.check(lib.scvae_constrained_poisson(_p(t), _ld(t), t.shape[0], _p(a), _ld(a), M, G,
                                             _p(count_sum), _p(row_const), _p(go), float(go_scalar),
                                             _p(da), _ld(da) if da is not None else 0, _p(logp),
                                             _p(lse), _stream()), "constrained_poisson")


def constrained_poisson_moments(a, lse, count_sum, B, G, RS, p_x_mean, p_x_stddev, stddev_of_mean):
    lib = _lib.load()
    _lib.check(lib.scvae_constrained_poisson_moments(_p(a), _ld(a), _p(lse), _p(count_sum), B, G, RS,
                                                     _p(p_x_mean), _p(p_x_stddev), _p(stddev_of_mean),
                                                     _ldo(p_x_mean, p_x_stddev, stddev_of_mean), _stream()),
               "constrained_poisson_moments")


def constrained_poisson_mixture_moments(a, lse, count_sum, B, G, RS, K_, y, p_x_mean, p_x_stddev,
                                        stddev_of_mean):
    """GMVAE: the constrained-Poisson moments marginalised over the K clusters."""
    lib = _lib.load()
    _lib.check(lib.scvae_constrained_poisson_mixture_moments(
        _p(a), _ld(a), _p(lse), _p(count_sum), B, G, RS, K_, _p(y), _ld(y), _p(p_x_mean),
        _p(p_x_stddev), _p(stddev_of_mean), _ldo(p_x_mean, p_x_stddev, stddev_of_mean), _stream()),
        "constrained_poisson_mixture_moments")


def likelihood_fwd(kind, t, a, head_stride, M, G, logp, row_const=None):
    _f32(t, a, logp)
    lib = _lib.load()
    _lib.check(lib.scvae_likelihood_fwd(kind, _p(t), _ld(t), t.shape[0], _p(a), _ld(a),
                                        head_stride, M, G, _p(row_const), _p(logp), _stream()),
               "likelihood_fwd")


def likelihood_bwd(kind, t, a, head_stride, M, G, da, logp=None, row_const=None, go=None,
                   go_scalar=1.0):
    _f32(t, a, da)
    lib = _lib.load()
    _lib.check(lib.scvae_likelihood_bwd(kind, _p(t), _ld(t), t.shape[0], _p(a), _ld(a),
                                        head_stride, M, G, _p(row_const), _p(go),
                                        float(go_scalar), _p(da), _ld(da), head_stride,
                                        _p(logp), _stream()), "likelihood_bwd")


def likelihood_moments(kind, a, head_stride, B, G, RS, K, y, p_x_mean, p_x_stddev,
                       stddev_of_mean):
    lib = _lib.load()
    ldo = _ldo(p_x_mean, p_x_stddev, stddev_of_mean)
    _lib.check(lib.scvae_likelihood_moments(kind, _p(a), _ld(a), head_stride, B, G, RS, K,
                                            _p(y), _ld(y) if y is not None else 0,
                                            _p(p_x_mean), _p(p_x_stddev), _p(stddev_of_mean),
                                            ldo, _stream()), "likelihood_moments")


def vae_bound(logp, kl_row, R, S, B, weight, out, go=None):
    lib = _lib.load()
    _lib.check(lib.scvae_vae_bound(_p(logp), _p(kl_row), R, S, B, float(weight), _p(out),
                                   _p(go), _stream()), "vae_bound")


def shadow(lo, hi, src_ld, cols, hi16, lo16=None, src_block_rows=1 << 40, dst_block_rows=1 << 40):
    """``scvae_shadow``: the fp32 block at flat offsets [lo, hi) (rows of src_ld floats) mirrored
    into the half matrix ``hi16`` (+ remainder ``lo16``)."""
    s = _lib.Shadow()
    s.lo, s.hi, s.src_ld, s.dst_ld = int(lo), int(hi), int(src_ld), _ld(hi16)
    s.src_block_rows, s.dst_block_rows = int(src_block_rows), int(dst_block_rows)
    s.hi16, s.lo16, s.cols = _p(hi16), _p(lo16), int(cols)
    return s


def adam_clip_ctas(n):
    return int(_lib.load().scvae_adam_clip_ctas(int(n)))


def adam_clip_step(param, grad, m, v, step, lr, beta1=0.9, beta2=0.999, epsilon=1e-8, clip=1.0,
                   grad_scale=1.0, scalars=None, shadows=None, advance_counter=None, advance_total=0):
    _f32(param, grad, m, v)
    lib = _lib.load()
    arr, n_sh = None, 0
    if shadows:
        n_sh = len(shadows)
        arr = (_lib.Shadow * n_sh)(*shadows)
    _lib.check(lib.scvae_adam_clip_step(_p(param), _p(grad), _p(m), _p(v), param.numel(),
                                        _p(step), float(lr), beta1, beta2, epsilon, clip,
                                        float(grad_scale), _p(scalars), arr, n_sh,
                                        _p(advance_counter), int(advance_total), _stream()),
               "adam_clip_step")


def dp_reduce_adam(world, rank, grad_ptrs, param_ptrs, flag_ptrs, m, v, n, step, lr, beta1=0.9,
                   beta2=0.999, epsilon=1e-8, clip=1.0, grad_scale=1.0, ctl=None, max_ctas=0,
                   scalars=None):
    """Fused reduce-scatter -> clip + Adam -> all-gather over peer memory (dp_exchange.cu).
    ``*_ptrs``: lists of ``world`` integer device addresses (peer-mapped), already offset to the
    start of the exchanged range; m, v: local tensors starting at the same offset."""
    import ctypes
    _f32(m, v)
    lib = _lib.load()
    arr = ctypes.c_void_p * world
    _lib.check(lib.scvae_dp_reduce_adam(world, rank, arr(*grad_ptrs), arr(*param_ptrs),
                                        arr(*flag_ptrs), _p(m), _p(v), int(n), _p(step), float(lr),
                                        beta1, beta2, epsilon, clip, float(grad_scale), _p(scalars),
                                        _p(ctl), int(max_ctas), _stream()), "dp_reduce_adam")


def step_advance(step):
    lib = _lib.load()
    _lib.check(lib.scvae_step_advance(_p(step), _stream()), "step_advance")


def col_mean(x, rows, cols, out):
    lib = _lib.load()
    _lib.check(lib.scvae_col_mean(_p(x), _ld(x), rows, cols, _p(out), _stream()), "col_mean")


def fill_normal(out, seed, offset=0, offset_dev=None):
    lib = _lib.load()
    _lib.check(lib.scvae_fill_normal(_p(out), out.numel(), int(seed), int(offset),
                                     _p(offset_dev), _stream()), "fill_normal")


# ---- Gaussian-mixture VAE pieces --------------------------------------------------------------
def group_offset_fwd(x, t, K_, B, H, y):
    lib = _lib.load()
    _lib.check(lib.scvae_group_offset_fwd(_p(x), _ld(x), _p(t), _ld(t), K_, B, H, _p(y), _ld(y),
                                          _stream()), "group_offset_fwd")


def group_offset_bwd(dy, K_, B, H, dx=None, dt=None, accumulate_dt=False):
    lib = _lib.load()
    _lib.check(lib.scvae_group_offset_bwd(_p(dy), _ld(dy), K_, B, H, _p(dx),
                                          _ld(dx) if dx is not None else 0, _p(dt),
                                          _ld(dt) if dt is not None else 0, int(accumulate_dt),
                                          _stream()), "group_offset_bwd")


def softmax_fwd(logits, B, K_, y, logy):
    lib = _lib.load()
    _lib.check(lib.scvae_softmax_fwd(_p(logits), _ld(logits), B, K_, _p(y), _p(logy), _stream()),
               "softmax_fwd")


def gmvae_latent_fwd(qh, pz, K_, B, L, RS, eps, z, klz, kl_elem=None):
    lib = _lib.load()
    _lib.check(lib.scvae_gmvae_latent_fwd(_p(qh), _ld(qh), _p(pz), K_, B, L, RS, _p(eps), _p(z),
                                          _ld(z), _p(klz), _p(kl_elem), _stream()),
               "gmvae_latent_fwd")


def gmvae_full_prior(pz, K_, L, pl):
    lib = _lib.load()
    _lib.check(lib.scvae_gmvae_full_prior(_p(pz), _ld(pz), K_, L, _p(pl), _stream()), "gmvae_full_prior")


def gmvae_latent_full_fwd(qh, pz, pl, K_, B, L, RS, eps, z, klz, w):
    lib = _lib.load()
    _lib.check(lib.scvae_gmvae_latent_full_fwd(_p(qh), _ld(qh), _p(pz), _ld(pz), _p(pl), K_, B, L, RS, _p(eps),
                                               _p(z), _ld(z), _p(klz), _p(w), _stream()),
               "gmvae_latent_full_fwd")


def gmvae_latent_full_bwd(qh, pz, pl, K_, B, L, RS, eps, dz, coef, w, cu, dqh, dpz):
    lib = _lib.load()
    _lib.check(lib.scvae_gmvae_latent_full_bwd(_p(qh), _ld(qh), _p(pz), _ld(pz), _p(pl), K_, B, L, RS, _p(eps),
                                               _p(dz), _ld(dz), _p(coef), _p(w), _p(cu), _p(dqh), _ld(dqh),
                                               _p(dpz), _ld(dpz), _stream()), "gmvae_latent_full_bwd")


def gmvae_full_covariance_mean(qh, K_, B, L, cov):
    lib = _lib.load()
    _lib.check(lib.scvae_gmvae_full_covariance_mean(_p(qh), _ld(qh), K_, B, L, _p(cov), _stream()),
               "gmvae_full_covariance_mean")


def gmvae_latent_bwd(qh, pz, K_, B, L, RS, eps, dz, coef, dqh, dpz):
    lib = _lib.load()
    _lib.check(lib.scvae_gmvae_latent_bwd(_p(qh), _ld(qh), _p(pz), K_, B, L, RS, _p(eps), _p(dz),
                                          _ld(dz), _p(coef), _p(dqh), _ld(dqh), _p(dpz),
                                          _stream()), "gmvae_latent_bwd")


def gmvae_row_coefficients(y, K_, RS, B, weight, go, coef):
    lib = _lib.load()
    _lib.check(lib.scvae_gmvae_row_coefficients(_p(y), K_, RS, B, float(weight), _p(go), _p(coef),
                                                _stream()), "gmvae_row_coefficients")


def gmvae_bound(y, logy, logp, klz, log_py, K_, RS, B, weight, free_nats_proportion, uniform_prior,
                out, dlogits, dpy_logits, ll_mean, klz_mean):
    lib = _lib.load()
    _lib.check(lib.scvae_gmvae_bound(_p(y), _p(logy), _p(logp), _p(klz), _p(log_py), K_, RS, B,
                                     float(weight), float(free_nats_proportion), int(uniform_prior),
                                     _p(out), _p(dlogits), _p(dpy_logits), _p(ll_mean),
                                     _p(klz_mean), _stream()), "gmvae_bound")


def gmvae_z_mean(qh, y, K_, B, L, z_mean):
    lib = _lib.load()
    _lib.check(lib.scvae_gmvae_z_mean(_p(qh), _ld(qh), _p(y), K_, B, L, _p(z_mean), _stream()),
               "gmvae_z_mean")


# ---- fused middle of the VAE step (csrc/mid_layers.cu) -------------------------------------------
def mid_layer(w=None, dw=None, beta=None, dbeta=None, moving_mean=None, moving_var=None, mean=None,
              rstd=None, y=None, n_in=0, k_in=0, n_out=0):
    """``scvae_mid_layer`` from tensors (2-D row-major ``w`` / ``dw`` / ``y``)."""
    l = _lib.MidLayer()
    for name, t in (("w", w), ("dw", dw), ("beta", beta), ("dbeta", dbeta),
                    ("moving_mean", moving_mean), ("moving_var", moving_var), ("mean", mean),
                    ("rstd", rstd), ("y", y)):
        setattr(l, name, _p(t))
    l.ldw = _ld(w) if w is not None else (_ld(dw) if dw is not None else 0)
    l.ldy = _ld(y) if y is not None else 0
    l.n_in, l.k_in, l.n_out = int(n_in), int(k_in), int(n_out)
    return l


def mid_desc():
    return _lib.MidDesc()


def vae_mid_workspace_floats(desc):
    import ctypes
    return int(_lib.load().scvae_vae_mid_workspace_floats(ctypes.byref(desc)))


def vae_mid_fwd(desc):
    import ctypes
    _lib.check(_lib.load().scvae_vae_mid_fwd(ctypes.byref(desc), _stream()), "vae_mid_fwd")


def vae_mid_bwd(desc):
    import ctypes
    _lib.check(_lib.load().scvae_vae_mid_bwd(ctypes.byref(desc), _stream()), "vae_mid_bwd")
