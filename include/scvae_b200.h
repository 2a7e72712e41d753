/*
 * scvae_b200.h -- C ABI of the B200-native scVAE hot path (libscvae_b200.so).
 *
 * The reference (scvae/scvae v2.1.4) has no FFI: its hot path is entered through
 * tf.Session.run (scvae/models/variational_autoencoder.py:1026-1029,
 * gaussian_mixture_variational_autoencoder.py:1109-1112) and every op below is a stock
 * TensorFlow / TensorFlow-Probability CPU op in that graph.  Each entry point cites the
 * reference graph fragment it replaces.  Abbreviations:
 *   VAE   = scvae/models/variational_autoencoder.py
 *   GMVAE = scvae/models/gaussian_mixture_variational_autoencoder.py
 *   MU    = scvae/models/utilities.py
 *   DU    = scvae/distributions/utilities.py
 *   ZI    = scvae/distributions/zero_inflated.py
 *
 * Conventions
 *   - plain pointers and sizes; every pointer is a DEVICE pointer unless marked [host];
 *   - caller owns all buffers, no hidden allocation, no hidden synchronisation;
 *   - `stream` is a cudaStream_t (CUstream) passed as void*;
 *   - return 0 on success, non-zero on error; the message is at scvae_last_error()
 *     (thread-local);
 *   - all matrices are row-major fp32 with an explicit leading dimension (in elements);
 *   - "augmented" activations: an activation matrix of logical width K is stored with
 *     width Kp >= K+1, column K == 1.0f and columns > K == 0, so that biases live in
 *     column K of the (out, Kp) weight matrices and need no separate kernel
 *     (fully_connected's BiasAdd, MU:53-59, and its gradient, fold into the GEMMs).
 */
#ifndef SCVAE_B200_H
#define SCVAE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SCVAE_B200_ABI_VERSION 1

/* Count likelihoods (DU:206-305). Head order = the reference's parameter order. */
#define SCVAE_LIK_POISSON 0 /* heads: log_lambda           DU:206-216 */
#define SCVAE_LIK_NB      1 /* heads: p, log_r             DU:266-281 */
#define SCVAE_LIK_ZIP     2 /* heads: pi, log_lambda       DU:247-264, ZI:180-199 */
#define SCVAE_LIK_ZINB    3 /* heads: pi, p, log_r         DU:283-305, ZI:180-199 */

/* Continuous / binary reconstruction distributions (DU:31-73, :125-245): their own row kernels
 * (scvae_continuous_likelihood / scvae_continuous_moments). */
#define SCVAE_LIK_GAUSSIAN          8  /* heads: mu, log_sigma                 DU:31-50 */
#define SCVAE_LIK_SOFTPLUS_GAUSSIAN 9  /* heads: mean, softplus_scale          DU:52-73 ("modified gaussian") */
#define SCVAE_LIK_LOG_NORMAL       10  /* heads: mean, variance                DU:125-140 */
#define SCVAE_LIK_GAMMA            11  /* heads: concentration, rate           DU:166-180 */
#define SCVAE_LIK_BERNOULLI        12  /* heads: logits                        DU:193-204 */
#define SCVAE_LIK_LOMAX            13  /* heads: log_concentration, log_scale  DU:230-245, lomax.py:177-247 */
#define SCVAE_LIK_EMG              14  /* heads: location, scale, rate         DU:142-164, exponentially_modified_normal.py:196-234 */

/* GEMM operand layouts. C is always (M, N) row-major. */
#define SCVAE_GEMM_NT 0 /* A (M,K) row-major, B (N,K) row-major : forward  Y = A W^T       */
#define SCVAE_GEMM_NN 1 /* A (M,K) row-major, B (K,N) row-major : dgrad    dA = dY W       */
#define SCVAE_GEMM_TN 2 /* A (K,M) row-major, B (K,N) row-major : wgrad    dW = dY^T A     */

int         scvae_abi_version(void);
const char *scvae_last_error(void);
int         scvae_num_heads(int kind);
/* Number of kernels this library has launched in the calling process (bench evidence). */
long long   scvae_launch_count(void);

/* ---- a1: minibatch gather  (VAE:994-998  x_train[idx].toarray()) --------------------
 * CSR (indptr int64 [n_rows+1], indices int32, values fp32) -> dense (B, ldx) fp32.
 * Row b of the output is CSR row rows[b] (rows == NULL: row b).  Columns [G, ldx) are
 * written as 1.0f at column G (if ldx > G) and 0 after it (augmented layout).
 * row_const (nullable, [B]) receives sum_g lgamma(1 + x[b,g]), the data-only constant of
 * every count log-likelihood.  rebase != 0: `indptr` holds absolute offsets of
 * a row slab whose nonzeros start at indices[0]/values[0] (offsets are taken relative to
 * indptr[0]) -- the streaming path ships one such slab per step from pinned host memory.
 * Further optional outputs (x, t16, x16 are each nullable, at least one must be given):
 * t16 ((B, ldt16) uint16, ldt16 % 8 == 0): counts clamped to 65535, targets of the fused
 * likelihood heads; x16 ((B, ldx16) fp16 augmented, ldx16 % 8 == 0, ldx16 > G): operand of the
 * fp16 tensor-core first layer.  Every output byte is written exactly once. */
int scvae_csr_densify(const int64_t *indptr, const int32_t *indices, const float *values,
                      const int64_t *rows, int B, int G, float *x, int64_t ldx,
                      float *row_const, int rebase, void *t16, int64_t ldt16, void *x16,
                      int64_t ldx16, void *stream);
/* The per-cell constant sum_g lgamma(1 + x) depends on the data only: computed ONCE for every
 * row of a CSR matrix when it is loaded (out [n_rows]; values fp32, or uint16 when
 * values_are_u16; absolute indptr), then gathered per minibatch (dst[b] = src[rows[b]], rows
 * == NULL: src[b]) instead of evaluating lgamma in every minibatch assembly. */
int scvae_csr_row_constants(const int64_t *indptr, const void *values, int values_are_u16,
                            int64_t n_rows, float *out, void *stream);
int scvae_gather_f32(const float *src, const int64_t *rows, int B, float *dst, void *stream);
/* Same with a compact CSR (uint16 column indices, G <= 65536, and uint16 integer counts):
 * 4 bytes per non-zero instead of 8 -- the format the streaming path ships over PCIe. */
int scvae_csr_densify_u16(const int64_t *indptr, const void *indices_u16, const void *values_u16,
                          const int64_t *rows, int B, int G, float *x, int64_t ldx,
                          float *row_const, int rebase, void *t16, int64_t ldt16, void *x16,
                          int64_t ldx16, void *stream);
/* Streamed form of the same gather (hotloop.PackedStream; replaces the reference's per-step
 * `x_train[batch_indices].toarray()` + feed_dict copy, VAE:985-1029, for matrices that stay in
 * host memory): the B rows of a minibatch arrive as ONE packed slab -- a single host -> device
 * copy per step, ~2 bytes per non-zero:
 *   int32 row_offset[B + 1] | float row_const[B] | (pad: strings start at
 *   scvae_packed_rows_offset(B)) | row strings, a row string (little endian, byte aligned) being
 *   u16 nesc | u16 nnz | u8 blocks[ceil(G / 255)] | (u8 index in block, u8 count)[nnz] |
 *   (u16 entry, u16 count)[nesc] | pad to a multiple of 16 bytes
 * blocks[k] = non-zeros of the row among the genes [255 k, 255 k + 255); a count byte of 255 is an
 * escape whose value is the escape-list entry with that entry position.  G <= 65280.
 * Outputs as the 16-bit outputs of scvae_csr_densify; row_const (nullable) receives the slab's copy.
 * scvae_pack_row_slab is the HOST routine that assembles such a slab (into pinned memory) from
 * the per-row strings encoded once per data set: store + row_off[n_rows + 1] (byte offsets),
 * order[rows] = the minibatch's rows; `threads` host threads copy the strings. */
int64_t scvae_packed_rows_offset(int B);
int scvae_csr_densify_packed(const void *slab, int B, int G, float *row_const, void *t16,
                             int64_t ldt16, void *x16, int64_t ldx16, void *stream);
/* The same assembly done by the GPU: `store` is PINNED HOST memory (read over PCIe through its
 * unified address), row_off / row_const_all / order are device arrays; two small kernels write
 * the slab into device memory.  No host work per step. */
int scvae_packed_pull(const void *store, const int64_t *row_off, const float *row_const_all,
                      const int64_t *order, int B, void *slab, int64_t slab_capacity, void *stream);
/* The same assembly done by the copy engine: the header goes through `header` (pinned, reusable
 * once the batch has executed), then ONE cudaMemcpyBatchAsync moves header + row strings from the
 * pinned store to the device slab.  Returns 2 when the runtime has no batched copies. */
int scvae_packed_copy_batch(const uint8_t *store, const int64_t *row_off, const float *row_const_all,
                            const int64_t *order, int rows, int64_t n_rows, uint8_t *header,
                            uint8_t *slab_dev, int64_t slab_capacity, int device, void *stream,
                            int64_t *bytes_out);
int scvae_pack_row_slab(const uint8_t *store, const int64_t *row_off, const float *row_const_all,
                        const int64_t *order, int rows, int64_t n_rows, uint8_t *dst,
                        int64_t dst_capacity, int threads, int64_t *bytes_out);
/* Dense fp32 counts -> uint16 (clamped), zero padded to ldt16 columns. */
int scvae_f32_to_u16(const float *x, int64_t ldx, int64_t rows, int G, void *t16, int64_t ldt16,
                     void *stream);

/* ---- a2: dense layers  (MU:38-76 fully_connected; its gradients) ----------------------
 * C[M,N] (+)= op(A) op(B), fp32 in / fp32 out.  `accumulate` != 0 adds into C.
 * scvae_gemm_f32 : exact fp32 FFMA kernel, any shape/alignment (small layers, parity).
 * scvae_gemm_tf32: tcgen05 tensor-core kernel (kind::tf32, TMA-staged, TMEM accumulators);
 *   requires 16-byte aligned bases and leading dimensions that are multiples of 4.
 *   `workspace` (nullable) of `workspace_bytes` enables deterministic split-K. */
int scvae_gemm_f32(int layout, int M, int N, int K, const float *A, int64_t lda,
                   const float *B, int64_t ldb, float *C, int64_t ldc, int accumulate,
                   void *stream);
int scvae_gemm_tf32(int layout, int M, int N, int K, const float *A, int64_t lda,
                    const float *B, int64_t ldb, float *C, int64_t ldc, int accumulate,
                    void *workspace, int64_t workspace_bytes, void *stream);
int64_t scvae_gemm_tf32_workspace_bytes(int layout, int M, int N, int K);
/* Same kernel with fp16 operands (kind::f16, fp32 accumulation, fp32 output scaled by alpha):
 * A, B are IEEE half matrices, leading dimensions (in elements) multiples of 8.  Used where a
 * (cells x genes)-sized operand is kept in 16 bits to halve its HBM traffic. */
int scvae_gemm_f16(int layout, int M, int N, int K, const void *A, int64_t lda, const void *B,
                   int64_t ldb, float *C, int64_t ldc, int accumulate, float alpha,
                   void *workspace, int64_t workspace_bytes, void *stream);
int64_t scvae_gemm_f16_workspace_bytes(int layout, int M, int N, int K);
/* fp16 product with one operand given as TWO fp16 matrices, X being the rounding remainder of the
 * fp32 original (scvae_f32_to_f16_split): which = 1: C = alpha (A + X) B, which = 2: C = alpha A (B + X);
 * X has the shape and layout of the operand it completes.  Both halves accumulate into the same
 * TMEM tile, so the product carries ~22 mantissa bits of that operand at fp16 HBM traffic for the
 * other one.  Used for the first encoder layer (weights split: its activations then match the
 * reference's fp32 `fully_connected`, MU:53-59, to ~1e-6) and its weight gradient (dY split: the
 * batch-norm backward makes that sum cancel heavily, so 11 bits are not enough). */
int scvae_gemm_f16_split(int layout, int M, int N, int K, const void *A, int64_t lda, const void *B,
                         int64_t ldb, const void *X, int64_t ldx, int which, float *C, int64_t ldc,
                         int accumulate, float alpha, void *workspace, int64_t workspace_bytes,
                         void *stream);
/* hi = fp16(scale * src), lo = fp16(scale * src - hi); zero padded to ldd columns (ldd % 8 == 0). */
int scvae_f32_to_f16_split(const float *src, int64_t lds, int64_t rows, int cols, void *hi, void *lo,
                           int64_t ldd, float scale, void *stream);
/* Bounds the persistent CTAs of the calling thread's subsequent tensor-core GEMM launches to
 * `max_ctas` (0 = one per SM, the default); returns the previous bound.  A caller that runs a
 * large HBM-bound product on a second stream uses it to leave SMs to the kernels it overlaps. */
int scvae_gemm_sm_limit(int max_ctas);
/* dst (rows, ldd) fp16 = scale * src (rows, lds) fp32 for the first `cols` columns, zero in
 * columns [cols, ldd): fp16 operand copies of fp32 master tensors. */
int scvae_f32_to_f16(const float *src, int64_t lds, int64_t rows, int cols, void *dst,
                     int64_t ldd, float scale, void *stream);

/* ---- f3: dropout on the input of a dense layer  (MU:45-50; one mask per site, VAE:2286,
 * :2487, :2516) ------------------------------------------------------------------------------
 * Inverted dropout on a bias-augmented operand: out = x * [noise < threshold] / keep on the
 * logical columns [0, n), which skip the physical column `skip_col` (the ones column when
 * decoder extras follow it; skip_col == n when it lies behind the masked range); every other
 * of the `width` stored columns is copied.  noise (rows, n) contiguous: standard-normal draws
 * (scvae_fill_normal) with threshold = the normal quantile of `keep`, or injected values.
 * Backward: dx (+)= dsrc * mask / keep on the masked columns (dsrc NULL: in place). */
int scvae_dropout_fwd(const float *x, int64_t ldx, int rows, int n, int skip_col,
                      const float *noise, float threshold, float keep, float *out, int64_t ldo,
                      int width, void *stream);
int scvae_dropout_bwd(float *dx, int64_t lddx, int rows, int n, int skip_col, const float *noise,
                      float threshold, float keep, const float *dsrc, int64_t ldds,
                      int accumulate, void *stream);

/* ---- a2: batch normalisation + ReLU  (MU:62-74; tf.contrib batch_norm center=True,
 * scale=False, epsilon 1e-3, decay 0.999) ---------------------------------------------
 * y (M, ldy) pre-activations, H logical columns; rows form `groups` consecutive groups of
 * M/groups rows with separate batch statistics (GMVAE: one BN op per cluster k,
 * GMVAE:2859-2877).  training != 0: batch stats (saved to save_mean/save_rstd
 * [groups*H]) and moving-average update (Bessel-corrected variance), applied group by
 * group in order; training == 0: moving stats.  relu != 0 applies max(.,0).
 * out (M, ldo) is written in augmented layout (col H = 1, cols > H = 0).
 * `scratch` must hold scvae_bn_scratch_floats(M, H, groups) floats. */
int64_t scvae_bn_scratch_floats(int M, int H, int groups);
int scvae_bn_act_fwd(const float *y, int64_t ldy, int M, int H, int groups,
                     const float *beta, float *moving_mean, float *moving_var,
                     int training, int update_moving, int relu, float *out, int64_t ldo,
                     float *save_mean, float *save_rstd, float *scratch, void *stream);
/* dout (M, lddo) gradient w.r.t. `out`; writes dy (M, lddy) and dbeta[H] (+= if
 * accumulate_dbeta).  Needs y, out and the saved statistics of the forward call. */
int scvae_bn_act_bwd(const float *dout, int64_t lddo, const float *y, int64_t ldy,
                     const float *out, int64_t ldo, int M, int H, int groups,
                     const float *save_mean, const float *save_rstd, int relu, float *dy,
                     int64_t lddy, float *dbeta, int accumulate_dbeta, float *scratch,
                     void *stream);
/* No-BN variant (minibatch_normalisation False): out = relu(y), augmented. */
int scvae_act_fwd(const float *y, int64_t ldy, int M, int H, int relu, float *out,
                  int64_t ldo, void *stream);
int scvae_act_bwd(const float *dout, int64_t lddo, const float *out, int64_t ldo, int M,
                  int H, int relu, float *dy, int64_t lddy, void *stream);

/* ---- a3/a6: Gaussian posterior, reparameterisation, analytic KL  (VAE:2243-2369,
 * VAE:2624-2656, DU:31-50) -------------------------------------------------------------
 * ph (B, ldph): columns [0,L) = mu pre-activation, [L,2L) = log_sigma pre-activation
 * (clipped to [-3,3]; ignored when unit_variance).  eps (RS*B, L) contiguous standard
 * normal noise, sample-major; deterministic != 0 -> z = mu, RS treated as 1.
 * z (RS*B, ldz) augmented.  kl_row[B] = sum_l KL(N(mu,sigma)||N(0,1)); kl_elem (nullable,
 * (B, L) contiguous) the per-neuron terms. */
int scvae_gaussian_latent_fwd(const float *ph, int64_t ldph, int B, int L, int RS,
                              const float *eps, int unit_variance, int deterministic,
                              float *z, int64_t ldz, float *kl_row, float *kl_elem,
                              void *stream);
/* dz (RS*B, lddz): gradient w.r.t. z.  kl_coef = d loss / d KL[b] (= warm_up*kl_weight/B).
 * dph (B, lddph) gradient w.r.t. the head pre-activations (clip has zero gradient outside
 * [-3,3]). */
int scvae_gaussian_latent_bwd(const float *ph, int64_t ldph, int B, int L, int RS,
                              const float *eps, int unit_variance, const float *dz,
                              int64_t lddz, float kl_coef, float *dph, int64_t lddph,
                              void *stream);
/* Sampled (non-analytical) KL term (VAE:2628-2640; the default of every VAE latent distribution
 * except `gaussian`, VAE:186-192): kl_rows[RS*B] (sample-major) = sum_l [log q(z|x) - log p(z)]
 * at z = mu + sigma eps, = sum_l (z^2/2 - eps^2/2 - log_sigma); kl_elem (nullable, (B, L)
 * contiguous) = the per-neuron terms averaged over the RS samples (column mean =
 * kl_divergence_neurons, VAE:2643-2646).  deterministic != 0: eps = 0, RS treated as 1. */
int scvae_gaussian_sampled_kl(const float *ph, int64_t ldph, int B, int L, int RS,
                              const float *eps, int unit_variance, int deterministic,
                              float *kl_rows, float *kl_elem, void *stream);
/* Backward of the sample and the sampled KL together.  c_m = d loss / d kl_rows[m] is
 * -weight * go[m] (go [RS*B] from scvae_vae_bound_rows) or, when go is NULL (R == 1),
 * coef_scalar = weight / (S B).  dph as in scvae_gaussian_latent_bwd. */
int scvae_gaussian_sampled_kl_bwd(const float *ph, int64_t ldph, int B, int L, int RS,
                                  const float *eps, int unit_variance, const float *dz,
                                  int64_t lddz, const float *go, float weight, float coef_scalar,
                                  float *dph, int64_t lddph, void *stream);
/* Decoder-input extras (VAE:2400-2441: `tf.concat([z, one_hot(batch_indices), count_sum])`):
 * for row m (cell b = m % B) of the latent sample matrix z (M, ldz) write the one-hot batch
 * index (batch_index [B] as float, nullable, n_batches columns) and/or the normalised count sum
 * (count_sum [B], nullable, one column) starting at column col0 (behind the ones column). */
int scvae_decoder_features(float *z, int64_t ldz, int M, int B, int col0, const float *batch_index,
                           int n_batches, const float *count_sum, void *stream);

/* ---- a5: count log-likelihood + reduction over genes  (VAE:2583-2590, DU:206-305,
 * ZI:194-199) --------------------------------------------------------------------------
 * a: head pre-activations (the FC outputs, before sigmoid / clip): head h of row m, gene g
 * at a[m*lda + h*head_stride + g].  t (t_rows, ldt) targets; row m uses t row m % t_rows
 * (replaces tf.tile, VAE:2564-2566).  row_const (nullable, [t_rows]) = sum_g lgamma(1+t).
 * logp[m] = sum_g log p(t[m,g] | theta[m,g]). */
int scvae_likelihood_fwd(int kind, const float *t, int64_t ldt, int t_rows, const float *a,
                         int64_t lda, int64_t head_stride, int M, int G,
                         const float *row_const, float *logp, void *stream);
/* Fused forward + backward: da (same layout as a: ldda, dhead_stride) = go[m] *
 * d logp[m] / d a  (go == NULL -> go_scalar for every row); logp nullable. */
int scvae_likelihood_bwd(int kind, const float *t, int64_t ldt, int t_rows, const float *a,
                         int64_t lda, int64_t head_stride, int M, int G,
                         const float *row_const, const float *go, float go_scalar,
                         float *da, int64_t ldda, int64_t dhead_stride, float *logp,
                         void *stream);

/* Piecewise-categorical ("categorised") likelihoods, the `-k` option (CAT:210-274, head P_K
 * VAE:2507-2532): log p(x) = log softmax(c)[min(x, k_max)] + [x >= k_max] log p_kind(x - k_max).
 * a (M, lda): the P heads of `kind` followed by the k_max + 1 class-logit heads (class-major),
 * all `head_stride` columns apart; da (nullable: forward only) has the same layout.  The
 * moments entry implements Categorised.mean / .variance (CAT:210-247) under VAE:2665-2713, for
 * the GMVAE (K > 1, rows ordered (k, sample, cell), y = q(y|x) [B, K]) marginalised over the
 * clusters as scvae_likelihood_moments. */
int scvae_piecewise_likelihood(int kind, int k_max, const float *t, int64_t ldt, int t_rows,
                               const float *a, int64_t lda, int64_t head_stride, int M, int G,
                               const float *go, float go_scalar, float *da, int64_t ldda,
                               int64_t dhead_stride, float *logp, void *stream);
int scvae_piecewise_moments(int kind, int k_max, const float *a, int64_t lda, int64_t head_stride,
                            int B, int G, int RS, int K, const float *y, int64_t ldy,
                            float *p_x_mean, float *p_x_stddev, float *stddev_of_mean,
                            int64_t ldo, void *stream);

/* Constrained Poisson (DU "constrained poisson", VAE:2492-2496): rate[m, g] = N[m % t_rows] *
 * clip(softmax_g(a[m, :]), tiny, 1); count_sum [t_rows] is the cell's count sum N (fed by the
 * reference as count_sum_parameter).  logp [M] (nullable) and/or da (M, ldda) = go[m] * d logp /
 * d a (NULL: forward only); lse [M] (nullable) receives the row log-sum-exp for the moments.
 * The moments entry gives p_x_mean / p_x_stddev / stddev_of_p_x_given_z_mean (VAE:2665-2713). */
int scvae_constrained_poisson(const float *t, int64_t ldt, int t_rows, const float *a, int64_t lda,
                              int M, int G, const float *count_sum, const float *row_const,
                              const float *go, float go_scalar, float *da, int64_t ldda,
                              float *logp, float *lse, void *stream);
int scvae_constrained_poisson_moments(const float *a, int64_t lda, const float *lse,
                                      const float *count_sum, int B, int G, int RS,
                                      float *p_x_mean, float *p_x_stddev, float *stddev_of_mean,
                                      int64_t ldo, void *stream);
/* GMVAE form (GMVAE:3170-3176, :3312-3386): a has K*RS*B rows ordered (k, sample, cell), lse one
 * entry per row, y (B, ldy) = q(y|x); moments marginalised over the clusters as in
 * scvae_likelihood_moments. */
int scvae_constrained_poisson_mixture_moments(const float *a, int64_t lda, const float *lse,
                                              const float *count_sum, int B, int G, int RS, int K,
                                              const float *y, int64_t ldy, float *p_x_mean,
                                              float *p_x_stddev, float *stddev_of_mean,
                                              int64_t ldo, void *stream);

/* ---- f3: continuous / binary reconstruction distributions -----------------------------------------
 * log p(t | a) summed over the genes of every (sample, cell) row and, when `da` is given, the
 * gradient go * d log p / d a, for kind = SCVAE_LIK_GAUSSIAN .. SCVAE_LIK_EMG.  `a` holds the P head
 * PRE-activations (head h at column h * head_stride); the kernel applies the head's activation and
 * clip (VAE:2466-2489: theta = clip(act(a), lo + tiny, hi - tiny), zero gradient outside the clip)
 * and the closed forms of TFP 0.7 Normal / LogNormal / Gamma / Bernoulli and of the reference's
 * Lomax / ExponentiallyModifiedNormal classes.  Targets tile over the rows (row m reads target row
 * m % t_rows).  scvae_continuous_moments: p_x_mean / p_x_stddev / stddev_of_p_x_given_z_mean as
 * scvae_likelihood_moments (Lomax: nan / inf where a moment does not exist, allow_nan_stats). */
int scvae_continuous_num_heads(int kind);
int scvae_continuous_likelihood(int kind, const float *t, int64_t ldt, int t_rows, const float *a,
                                int64_t lda, int64_t head_stride, int M, int G, const float *go,
                                float go_scalar, float *da, int64_t ldda, int64_t dhead_stride,
                                float *logp, void *stream);
int scvae_continuous_moments(int kind, const float *a, int64_t lda, int64_t head_stride, int B, int G,
                             int RS, int K, const float *y, int64_t ldy, float *p_x_mean,
                             float *p_x_stddev, float *stddev_of_mean, int64_t ldo, void *stream);

/* ---- a4 + a5 fused: likelihood heads without the (cells x P*genes) round trip ----------------
 * One kernel computes a = d W^T (tcgen05, fp16 operands), log p(t | a) summed over genes, its
 * gradient da and the decoder gradient dd = da W; only da (fp16, scaled by `scale`) is written
 * to HBM, for the weight-gradient product dW = da^T d (scvae_gemm_f16 with alpha = 1/scale).
 *   d16  (M, 128) fp16: decoder output, augmented, zero padded to 128 columns (n_in + 1 <= 128)
 *   w16  (P*head_stride, 128) fp16: head weights, head h at rows [h*head_stride, +G), zero
 *        padded; head_stride % 64 == 0
 *   t16  (t_rows, ldt) targets, uint16 (t_is_half == 0) or fp16 (t_is_half != 0, exact for
 *        counts <= 2048); row m uses t row m % t_rows; t_rows % 128 == 0 or == M
 *   da16 (M, P*head_stride) fp16 out; dd (M, lddd) fp32 out (first dd_cols columns);
 *   logp [M] out; workspace: scvae_heads_fused_workspace_floats(M, G) floats. */
int64_t scvae_heads_fused_workspace_floats(int M, int G);
/* Forward only (the per-epoch evaluation passes VAE:1092-1150 and `evaluate` when no per-gene
 * moments are requested): a = d W^T and log p(t | a) summed over genes; nothing but logp [M]
 * leaves the chip.  Same operand layouts as below. */
int scvae_heads_fused_fwd(int kind, const void *d16, const void *w16, int64_t head_stride,
                          const void *t16, int64_t ldt, int t_is_half, int t_rows, int M, int G,
                          const float *row_const, float *logp, float *workspace, void *stream);
int scvae_heads_fused_bwd(int kind, const void *d16, const void *w16, int64_t head_stride,
                          const void *t16, int64_t ldt, int t_is_half, int t_rows, int M, int G,
                          const float *row_const, const float *go, float go_scalar, float scale,
                          void *da16, float *dd, int64_t lddd, int dd_cols, float *logp,
                          float *workspace, void *stream);

/* ---- a7: evaluate-only moments  (VAE:2534-2552, VAE:2665-2713, ZI:180-192;
 * GMVAE:3312-3386) ---------------------------------------------------------------------
 * a has K*RS*B rows ordered (k, rs, b).  y (nullable, (B, ldy)) cluster weights q(y|x)
 * (K == 1 and y == NULL for the VAE).  Outputs (B, ldo), any nullable:
 * p_x_mean, p_x_stddev, stddev_of_p_x_given_z_mean. */
int scvae_likelihood_moments(int kind, const float *a, int64_t lda, int64_t head_stride,
                             int B, int G, int RS, int K, const float *y, int64_t ldy,
                             float *p_x_mean, float *p_x_stddev, float *stddev_of_mean,
                             int64_t ldo, void *stream);

/* ---- a6: VAE bound  (VAE:2715-2734, MU:129-137) -----------------------------------------
 * logp[R*S*B] (r, s, b), kl_row[B].  out[4] = {lower_bound, lower_bound_weighted,
 * reconstruction_error, kl_divergence}.  go (nullable, [R*S*B]) = d(-lower_bound_weighted)
 * / d logp = -softmax_r(logp - w kl)/(S B).  weight = warm_up_weight * kl_weight. */
int scvae_vae_bound(const float *logp, const float *kl_row, int R, int S, int B,
                    float weight, float *out, float *go, void *stream);

/* The same bound with one KL value per (r, s, b) row (sampled KL, VAE:2656, :2715-2734):
 * kl_rows[R*S*B]; out[3] = mean over all rows (= sum_l kl_divergence_neurons). */
int scvae_vae_bound_rows(const float *logp, const float *kl_rows, int R, int S, int B,
                         float weight, float *out, float *go, void *stream);

/* ---- a8: optimiser  (VAE:2736-2770) elementwise clip to [-clip, clip] + TF Adam -------
 * One fused pass over the flat parameter buffer.  `step` [device, int64] is read (t =
 * *step + 1) for the bias correction lr_t = lr sqrt(1-b2^t)/(1-b1^t); the caller advances
 * it with scvae_step_advance (kept separate so the pair is CUDA-graph capturable).
 * grad_scale multiplies the gradient before clipping (data-parallel mean).
 * scalars (nullable, device {learning rate factor, warm-up weight}): lr is multiplied by
 * scalars[0], so a captured step serves any learning rate.
 * shadows (nullable, [host], n_shadows <= SCVAE_MAX_SHADOWS): blocks of the parameter range whose
 * fp16 operand copies (+ rounding remainders, see scvae_gemm_f16_split) are rewritten in the same
 * pass: flat offsets [lo, hi) relative to `param` hold rows of src_ld floats; row r goes to row
 * (r / src_block_rows) * dst_block_rows + r % src_block_rows of the (.., dst_ld) half matrix,
 * columns < cols copied, the rest of a touched 4-column group zeroed.
 * advance_counter (nullable, device int, zero-initialised) / advance_total: *step += 1 once the
 * last of `advance_total` CTAs sharing the counter finishes (scvae_adam_clip_ctas(n) CTAs per
 * launch) -- replaces scvae_step_advance when every optimiser launch of a step takes part. */
#define SCVAE_MAX_SHADOWS 4
typedef struct scvae_shadow {
    int64_t lo, hi, src_ld, dst_ld;
    int64_t src_block_rows, dst_block_rows;
    void *hi16, *lo16;
    int cols, reserved;
} scvae_shadow;
int scvae_adam_clip_step(float *param, const float *grad, float *m, float *v, int64_t n,
                         int64_t *step, float lr, float beta1, float beta2,
                         float epsilon, float clip, float grad_scale, const float *scalars,
                         const scvae_shadow *shadows /* [host] */, int n_shadows,
                         int *advance_counter, int advance_total, void *stream);
int scvae_adam_clip_ctas(int64_t n);
int scvae_step_advance(int64_t *step, void *stream);

/* ---- e: data-parallel exchange fused with the optimiser ----------------------
 * Replaces  all-reduce(grad) ; scvae_adam_clip_step(all parameters)  on W ranks of one NVLink
 * domain by ONE kernel per rank: reduce-scatter of the flat gradient range over peer memory
 * (P2P loads, fixed summation order), clip + TF-Adam on this rank's 1/W slice (only that slice
 * of m / v is touched), all-gather of the updated parameters (P2P stores).  grad_ptrs /
 * param_ptrs / flag_ptrs are HOST arrays of `world` peer-mapped device pointers to the start of
 * the range in every rank's gradient / parameter buffer and to every rank's flag block (32
 * zero-initialised uint32 per exchange channel).  ctl: 4 zero-initialised device uint32 owned
 * by the channel ([2] != 0 after a peer timed out).  Two flag barriers per launch; no host
 * synchronisation; CUDA-graph capturable.  n % 4 == 0.  scalars: as in scvae_adam_clip_step. */
int scvae_dp_reduce_adam(int world, int rank, const void *const *grad_ptrs,
                         void *const *param_ptrs, void *const *flag_ptrs, float *m, float *v,
                         int64_t n, const int64_t *step, float lr, float beta1, float beta2,
                         float epsilon, float clip, float grad_scale, const float *scalars, void *ctl,
                         int max_ctas, void *stream);

/* ---- a9/a10: Gaussian-mixture VAE pieces  (GMVAE:2788-3434) ------------------------------
 * The K cluster passes are K consecutive row groups of one tall matrix, rows ordered
 * (k, sample, cell); dense layers / batch norm (groups = K) / likelihood are shared with the VAE.
 *
 * group_offset: y[k*B + b, :H] = x[b, :H] + t[k, :H]  -- the one-hot concat [x, e_k] of
 * GMVAE:2942-2947 (t = rows of the weight block acting on e_k; x W_x is computed once).
 * Backward: dx[b] = sum_k dy[k*B+b] (nullable), dt[k] = sum_b dy[k*B+b] (nullable, += if
 * accumulate_dt).  With B == 1 it is FC(one_hot_k) = W[k] + b of GMVAE:3009-3048. */
int scvae_group_offset_fwd(const float *x, int64_t ldx, const float *t, int64_t ldt, int K,
                           int B, int H, float *y, int64_t ldy, void *stream);
int scvae_group_offset_bwd(const float *dy, int64_t lddy, int K, int B, int H, float *dx,
                           int64_t lddx, float *dt, int64_t lddt, int accumulate_dt,
                           void *stream);
/* q(y|x) = Cat(logits)  (GMVAE:3050-3092): y, logy (B, K) contiguous. */
int scvae_softmax_fwd(const float *logits, int64_t ldl, int B, int K, float *y, float *logy,
                      void *stream);
/* q(z|x,y=k) = N(mean, sqrt(softplus(s))), z = mean + scale*eps, sampled
 * KL_z = sum_l log q(z) - log p(z|y=k)  (GMVAE:2936-3048, 3270-3289; DU:52-73).
 * qh (K*B, ldq) = [mean | s]; pz (K, 2L) contiguous = [mean | s] of p(z|y=k);
 * eps (K*RS*B, L); z (K*RS*B, ldz) augmented; klz [K*RS*B]; kl_elem nullable (K*RS*B, L). */
int scvae_gmvae_latent_fwd(const float *qh, int64_t ldq, const float *pz, int K, int B, int L,
                           int RS, const float *eps, float *z, int64_t ldz, float *klz,
                           float *kl_elem, void *stream);
/* coef [K*RS*B] = d loss / d klz; dz decoder gradient; dqh (K*B, lddq), dpz (K, 2L). */
int scvae_gmvae_latent_bwd(const float *qh, int64_t ldq, const float *pz, int K, int B, int L,
                           int RS, const float *eps, const float *dz, int64_t lddz,
                           const float *coef, float *dqh, int64_t lddq, float *dpz,
                           void *stream);
/* The same block for the FULL-COVARIANCE mixture (`-q "full-covariance gaussian mixture"`): q and p
 * are multivariate Gaussians with a lower-triangular scale matrix S (DU:75-93: L locations +
 * T = L (L + 1) / 2 scales, softplus clipped at tiny from below, laid out by
 * tfp.distributions.fill_triangular; multivariate_normal.py:90-150; call sites GMVAE:2958-3048,
 * :3270-3289).  qh (K*B, ldq) = [locations | raw scales] (L + T columns), pz (K, ldp) likewise;
 * z = loc_q + S_q eps; klz = log q(z) - log p(z|y=k) with one triangular solve per row.
 * scvae_gmvae_full_prior expands the K activated prior scale matrices pl (K, L, L) once per step;
 * the forward keeps w (K*RS*B, L) = S_p^-1 (z - loc_p) for the backward, whose scratch cu has the
 * same shape; dqh (K*B, lddq), dpz (K, lddp) receive the gradients w.r.t. the pre-activations.
 * scvae_gmvae_full_covariance_mean: cov (K, L, L) = mean over the cells of S_q S_q^T
 * (q_z_covariances; its diagonal = q_z_variances, GMVAE:2881-2893).  L <= 128. */
int scvae_gmvae_full_prior(const float *pz, int64_t ldp, int K, int L, float *pl, void *stream);
int scvae_gmvae_latent_full_fwd(const float *qh, int64_t ldq, const float *pz, int64_t ldp,
                                const float *pl, int K, int B, int L, int RS, const float *eps,
                                float *z, int64_t ldz, float *klz, float *w, void *stream);
int scvae_gmvae_latent_full_bwd(const float *qh, int64_t ldq, const float *pz, int64_t ldp,
                                const float *pl, int K, int B, int L, int RS, const float *eps,
                                const float *dz, int64_t lddz, const float *coef, const float *w,
                                float *cu, float *dqh, int64_t lddq, float *dpz, int64_t lddp,
                                void *stream);
int scvae_gmvae_full_covariance_mean(const float *qh, int64_t ldq, int K, int B, int L, float *cov,
                                     void *stream);
/* go = -y[b,k]/(B RS) (gradient of the loss w.r.t. each row's log-likelihood),
 * coef = weight * y[b,k]/(B RS) (w.r.t. each row's KL_z); both [K*RS*B]. */
int scvae_gmvae_row_coefficients(const float *y, int K, int RS, int B, float weight, float *go,
                                 float *coef, void *stream);
/* y-marginalised bound (GMVAE:3242-3410). logp, klz [K*RS*B]; log_py [K] (log prior
 * probabilities; -log K for the uniform prior). out[6] = {lower_bound, lower_bound_weighted,
 * reconstruction_error, kl_divergence_z, kl_divergence_y, kl_divergence_y after free nats}.
 * dlogits (nullable, (B,K)), dpy_logits (nullable, [K]); ll_mean, klz_mean (K, B) scratch /
 * outputs (sample means, not y-weighted). */
int scvae_gmvae_bound(const float *y, const float *logy, const float *logp, const float *klz,
                      const float *log_py, int K, int RS, int B, float weight,
                      float free_nats_proportion, int uniform_prior, float *out, float *dlogits,
                      float *dpy_logits, float *ll_mean, float *klz_mean, void *stream);
/* z_mean[b] = sum_k y[b,k] mean_k[b]  (GMVAE:2896-2899); z_mean (B, L) contiguous. */
int scvae_gmvae_z_mean(const float *qh, int64_t ldq, const float *y, int K, int B, int L,
                       float *z_mean, void *stream);

/* ---- a2 + a3 + a6 fused: the (cells x ~100) middle of a VAE training step --------------------
 * Everything between the two gene-axis products of a step -- per `dense_layer` (MU:53-74) the
 * FC with batch norm and ReLU behind it, the posterior heads with clip / reparameterised sample /
 * analytic KL behind them (VAE:2280-2369, :2624-2627), the decoder layers -- as ONE persistent
 * kernel per direction: a CTA owns <= 64 cells, keeps their activations in shared memory and runs
 * every small product in exact fp32 with its normalisation / activation / sample as the epilogue;
 * batch statistics and weight-gradient partials cross CTAs through `workspace` and a grid barrier
 * and are folded in a fixed order (deterministic).  Requirements: hidden widths < 128, latent
 * size <= 128, one sample per cell (R = S = 1), B <= 64 * (number of SMs).
 *
 * scvae_vae_mid_fwd:  enc[0].y partials (the tensor-core product x W1^T, `y1_nsplit` split-K
 *   slices `y1_slice` floats apart, scaled by y1_alpha) -> ... -> d16, the fp16 augmented operand
 *   of scvae_heads_fused_*.  Writes every layer's pre-activations `y` and batch statistics
 *   `mean` / `rstd` (training), ph = [mu | raw log_sigma], z (augmented, with the decoder-input
 *   extras of VAE:2400-2441), kl_row, optionally kl_elem, and the noise when generate_eps (same
 *   Philox stream as scvae_fill_normal(seed, offset + *offset_dev)).
 * scvae_vae_mid_bwd:  dd_parts / logp_parts = the gene-range partials of scvae_heads_fused_bwd's
 *   workspace -> logp (with the first-order correction for the fp16 rounding of d16), bound[4] =
 *   {lower_bound, lower_bound_weighted, reconstruction_error, kl_divergence} (VAE:2715-2734, R = 1),
 *   dw / dbeta of every layer but the first encoder layer, and dy1_16 (+ dy1_16_lo, nullable: the
 *   rounding remainder, see scvae_gemm_f16_split) = fp16(dy1_scale * dY1) for that layer's
 *   weight-gradient product.  enc[0].dw (nullable; ldw = its row pitch): only its bias column n_in
 *   is written, = the column sums of dY1 in fp32 -- for callers that form the gene columns from a
 *   single fp16 dy1_16, whose rounding would otherwise be all there is in that (behind a batch
 *   norm: mathematically zero) gradient.  kl weight = kl_weight * scalars[1] (scalars nullable,
 *   device: {learning rate, warm-up weight}).
 * A barrier wait that exceeds ~2 s sets *error (device int) instead of hanging. */
#define SCVAE_MID_MAX_LAYERS 4
typedef struct scvae_mid_layer {
    const float *w;        /* (n_out, ldw): bias in column n_in, decoder-input extras behind it */
    float *dw;             /* gradient of w (backward) */
    const float *beta;     /* batch-norm offset [n_out]; NULL = no batch norm */
    float *dbeta;
    float *moving_mean, *moving_var;
    float *mean, *rstd;    /* batch statistics [n_out]: written forward, read backward */
    float *y;              /* pre-activations (B, ldy): written forward, read backward */
    int64_t ldw, ldy;
    int n_in, k_in, n_out; /* k_in = n_in + 1 + extras = reduction length of the forward product */
    int reserved;
} scvae_mid_layer;
typedef struct scvae_mid_desc {
    int B, L, n_enc, n_dec;
    int training, update_moving, deterministic, rows_per_cta;
    scvae_mid_layer enc[SCVAE_MID_MAX_LAYERS], post, dec[SCVAE_MID_MAX_LAYERS];
    const float *y1_parts; int64_t y1_ld, y1_slice; int y1_nsplit; float y1_alpha;
    float *ph; int64_t ldph;
    float *eps; int generate_eps; int reserved0; uint64_t seed, offset; const int64_t *offset_dev;
    float *z; int64_t ldz;
    float *kl_row, *kl_elem;
    const float *batch_index; const float *count_sum; int n_batches; int reserved1;
    void *d16; int64_t ldd16;
    float *h_last; int64_t ldh_last;           /* optional fp32 copy of the last decoder activation */
    /* backward */
    const float *dd_parts; int64_t dd_ld, dd_slice; int dd_nsplit; int logp_nsplit;
    const float *logp_parts; int64_t logp_slice; const float *row_const;
    float *logp; float *bound;
    void *dy1_16; void *dy1_16_lo; int64_t lddy1; float *dy1; int64_t lddy1_f32;
    float go_scalar, dy1_scale, kl_weight; int reserved2; const float *scalars;
    float *workspace; int64_t workspace_floats; uint32_t *barrier; int *error;
    long long *timeline;    /* development aid (nullable): [CTA][32] %globaltimer stamps of the phases */
} scvae_mid_desc;
int64_t scvae_vae_mid_workspace_floats(const scvae_mid_desc *desc /* [host] */);
int scvae_vae_mid_fwd(const scvae_mid_desc *desc /* [host] */, void *stream);
int scvae_vae_mid_bwd(const scvae_mid_desc *desc /* [host] */, void *stream);

/* ---- small helpers used by the shells ------------------------------------------------- */
/* out[c] = (1/rows) * sum_r x[r, c]  (kl_divergence_neurons, VAE:2643-2646). */
int scvae_col_mean(const float *x, int64_t ldx, int rows, int cols, float *out,
                   void *stream);
/* Philox-4x32-10 standard-normal fill (tf.random_normal stand-in, VAE:2363).  The Philox
 * offset is `4 * (offset + *offset_dev)` outputs, i.e. one Philox block per unit, so calls with
 * different offsets draw from disjoint blocks (offset_dev nullable, device int64: normally the
 * optimiser step counter, which keeps the launch CUDA-graph capturable). */
int scvae_fill_normal(float *out, int64_t n, uint64_t seed, uint64_t offset,
                      const int64_t *offset_dev, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* SCVAE_B200_H */
