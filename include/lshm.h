/*
 * liblshm_sm100 — C ABI of the B200-native LSHM deep-K-harmonic hot path.
 *
 * The reference (SarodYatawatta/LSHM) is pure Python/PyTorch and has NO native / FFI
 * boundary of its own.  Each entry point below therefore replaces a
 * *Python-level* interface of the reference, cited as file:line relative to
 * /root/reference.  INTEGRATION.md shows the ctypes stub a reference maintainer adds.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer owned by the caller (outputs pre-allocated);
 *    the library keeps no persistent device state;
 *  - `stream` is a cudaStream_t passed as void*; every call only enqueues work on it;
 *  - return 0 on success, negative lshm_status on error; lshm_last_error() gives a
 *    thread-local message.  NaN/Inf in data is not an error (it propagates);
 *  - all tensors are dense fp32 unless stated; "ns"/"ld" arguments are strides in
 *    ELEMENTS between consecutive samples / rows.
 *
 * Geometry shared by the conv entry points ("small" map S, "big" map B, weight W):
 *    2-D (k4 s2 p1):  S [N,A,h,w]   B [N,Bc,2h,2w]   W [A,Bc,4,4]
 *                     big (y,x) = (2*oy-1+ky, 2*ox-1+kx)
 *    1-D (k4 s4 pad): S [N,A,l]     B [N,Bc,4l]      W [A,Bc,4]
 *                     big pos   = 4*j - pad + t      (pad=1 Conv1d, pad=0 ConvTranspose1d)
 *  Conv:          S = act(W * B + bias[A])        ("down")
 *  ConvTranspose: B = act(W^T * S + bias[Bc])     ("up")
 *  and the backward of one is the other, so three kernels serve six layer types; the weight gradient
 *  ("wgrad") and the data gradient of the last two transposed convs also exist as ONE launch
 *  (lshm_tconv_bwd*: their output gradient, the layer's largest tensor, is read once).
 *  act / epilogue codes: LSHM_EPI_*.
 */
#ifndef LSHM_H_
#define LSHM_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define LSHM_API __attribute__((visibility("default")))
#else
#define LSHM_API
#endif

typedef void* lshm_stream_t;

enum lshm_status {
  LSHM_OK = 0,
  LSHM_ERR_ARG = -1,      /* bad shape / null pointer / unsupported size */
  LSHM_ERR_CUDA = -2,     /* CUDA runtime reported an error at launch */
  LSHM_ERR_UNSUPPORTED = -3
};

enum lshm_epilogue {
  LSHM_EPI_NONE = 0,      /* out = acc (+bias if given) */
  LSHM_EPI_ELU = 1,       /* out = ELU(acc + bias), alpha = 1 (F.elu, src/lofar_models.py:73) */
  LSHM_EPI_DELU = 2       /* out = acc * ELU'(aux), aux = post-ELU activation at the same
                             location:  ELU'(z) = aux>0 ? 1 : aux+1  (backward through F.elu) */
};

LSHM_API const char* lshm_last_error(void);
LSHM_API int lshm_version(void);
/* SM count / compute capability of the current device (host query, no launch). */
LSHM_API int lshm_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* ------------------------------------------------------------------ loader ------
 * Replaces get_data_minibatch / get_data_for_baseline bodies,
 * src/lofar_tools.py:113-141 (int8 x scale), :157-173 (patchify, patch-major rows
 * n=(ci*py+cj)*nb+k), :187 / :333 (clamp), :190-193 / :336-338 (global z-score).
 * vis [nbase,T,F,4,2] int8, scale [nbase,F,4] fp32, sel [nb] int32 baseline ids.
 * C in {4,8}: C=8 -> channel 2*pol+ri ; C=4 -> pols 0 and 3.
 * y [nb*px*py, C, P, P]; zero padded when T<P or F<P.  stats (double[2], may be NULL)
 * receives sum and sum of squares of every written element (after clamp); it must be
 * zeroed by the caller. */
LSHM_API int lshm_patchify_scale_i8(const int8_t* vis, const float* scale, const int32_t* sel,
                           int nb, int T, int F, int C, int P, float clamp,
                           float* y, double* stats, lshm_stream_t stream);
/* y = (y - mean) / std, UNBIASED std (torch.Tensor.std), from stats of n elements. */
LSHM_API int lshm_normalise(float* y, int64_t n, const double* stats, lshm_stream_t stream);
/* Same for a SHARD of the minibatch: y holds n of the n_stats elements the (all-reduced) stats describe, so
 * every data-parallel rank z-scores with the statistics of the whole minibatch. */
LSHM_API int lshm_normalise_n(float* y, int64_t n, const double* stats, int64_t n_stats, lshm_stream_t stream);

/* ------------------------------------------------------------- FFT features -----
 * Replaces Demo.ipynb:169-174 + torch_fftshift (src/lofar_tools.py:24-30):
 * F = fftshift(fft2_ortho(x - xhat)); mode LSHM_FFT_REIM: out[:, :C] = Re F, out[:, C:] = Im F, both clamped to
 * +-clamp (the reference feature); mode LSHM_FFT_MAGPHASE: out[:, :C] = min(|F|, clamp), out[:, C:] = arg F
 * in (-pi, pi] (torch.abs / torch.angle of the same F).  x, xhat [N,C,128,128] (xhat may be NULL),
 * out [N,2C,128,128]; x and out 16-byte aligned. */
enum lshm_fft_mode { LSHM_FFT_REIM = 0, LSHM_FFT_MAGPHASE = 1 };
LSHM_API int lshm_fft2_features(const float* x, const float* xhat, float* out,
                       int64_t N, int C, float clamp, int mode, lshm_stream_t stream);
/* = lshm_fft2_features(..., LSHM_FFT_REIM, ...) */
LSHM_API int lshm_fft2_reim_shift_clamp(const float* x, const float* xhat, float* out,
                               int64_t N, int C, float clamp, lshm_stream_t stream);

/* -------------------------------------------------------------- autoencoders ----
 * uv [N,2], scales [H] -> out [N,4H] = [sin(s_h*u), sin(s_h*v) ... | cos ...]
 * (torch.kron + cat(sin,cos), src/lofar_models.py:60-62). */
LSHM_API int lshm_uv_harmonics(const float* uv, const float* scales, int64_t N, int H, float* out,
                      lshm_stream_t stream);

/* Weight "images" for the tensor-core conv kernels: the fp32 weight W[A,Bc,4(,4)] split into
 * bf16 hi/lo halves and permuted into the shared-memory operand layout of each kernel, so a CTA
 * fetches its weight tile with one bulk copy.  Must be re-made whenever W changes (once per
 * closure evaluation).  which: 0 = image for lshm_down*, 1 = image for lshm_up*. */
LSHM_API int lshm_conv_image_bytes(int dim, int A, int Bc, int which, int64_t* bytes);
LSHM_API int lshm_conv_prep(const float* w, int dim, int A, int Bc, void* down_img, void* up_img,
                   lshm_stream_t stream);
/* Batched form: the caller fills one 8 x int64 record per image on the HOST with
 * lshm_conv_prep_record, uploads the table once (weight and image addresses are stable), and
 * lshm_conv_prep_batch re-makes all n images with a single launch. */
LSHM_API int lshm_conv_prep_record(const float* w, int dim, int A, int Bc, int which, void* img, int64_t* record);
LSHM_API int lshm_conv_prep_batch(const int64_t* table, int n, lshm_stream_t stream);

/* Conv2d(k4,s2,p1) forward (src/lofar_models.py:73-78) and ConvTranspose2d dgrad.
 * wimg = "down" image of W (lshm_conv_prep). */
LSHM_API int lshm_down2d(const float* big, int64_t big_ns, const void* wimg, const float* bias,
                const float* aux, int64_t aux_ns, float* small_, int64_t small_ns,
                int64_t N, int A, int Bc, int h, int w_, int epilogue, lshm_stream_t stream);
/* ConvTranspose2d(k4,s2,p1) forward (src/lofar_models.py:93-98) and Conv2d dgrad.
 * wimg = "up" image of W (lshm_conv_prep). */
LSHM_API int lshm_up2d(const float* small_, int64_t small_ns, const void* wimg, const float* bias,
              const float* aux, int64_t aux_ns, float* big, int64_t big_ns,
              int64_t N, int A, int Bc, int h, int w_, int epilogue, lshm_stream_t stream);
/* dW[A,Bc,4,4] = sum_n small (x) big ; written, not accumulated. */
LSHM_API int lshm_wgrad2d(const float* small_, int64_t small_ns, const float* big, int64_t big_ns,
                 float* dw, int64_t N, int A, int Bc, int h, int w_, lshm_stream_t stream);

/* Conv1d(k4,s4,p1) forward (src/lofar_models.py:158-163), ConvTranspose1d dgrad (pad=0). */
LSHM_API int lshm_down1d(const float* big, int64_t big_ns, const void* wimg, const float* bias,
                const float* aux, int64_t aux_ns, float* small_, int64_t small_ns,
                int64_t N, int A, int Bc, int l, int pad, int epilogue, lshm_stream_t stream);
/* ConvTranspose1d(k4,s4,p0) forward (src/lofar_models.py:178-183), Conv1d dgrad (pad=1). */
LSHM_API int lshm_up1d(const float* small_, int64_t small_ns, const void* wimg, const float* bias,
              const float* aux, int64_t aux_ns, float* big, int64_t big_ns,
              int64_t N, int A, int Bc, int l, int pad, int epilogue, lshm_stream_t stream);
LSHM_API int lshm_wgrad1d(const float* small_, int64_t small_ns, const float* big, int64_t big_ns,
                 float* dw, int64_t N, int A, int Bc, int l, int pad, lshm_stream_t stream);
/* ---- operand planes -----------------------------------------------------------------------------
 * An input-sized tensor ("big" map B of the geometry above) stored the way the tensor-core kernels consume
 * it: the space-to-depth operand Z[q, c = 4*b + sub] (2-D: q over the (h+1) x (w+1) grid of 2x2 pixel blocks
 * at rows 2by-1, 2by / columns 2bx-1, 2bx incl. the zero halo, sub = 2*sy + sx; 1-D: q = window j,
 * sub = t, sample 4j - pad + t) split into bf16 hi = bf16(v), lo = bf16(v - hi):
 *     planes = [half hi|lo][chunk c/8][position q][8 x bf16]          (lshm_planes_bytes bytes)
 * lshm_down*_planes / lshm_wgrad*_planes fetch their tiles from it with tensor-TMA box loads
 * (cp.async.bulk.tensor.3d) instead of gathering + converting fp32 in producer warps.  Supported for the
 * first layers (A <= 16 output channels, Bc = 8 / Bc <= 16 input channels), whose inputs are what the
 * loader / the loss kernels produce. */
LSHM_API int lshm_planes_bytes(int dim, int64_t N, int Bc, int h, int w_or_l, int64_t* bytes);
/* planes <- fp32 big map [N,Bc,2h,2w] (sample stride big_ns) / [N,Bc,4l] with the 1-D padding baked in. */
LSHM_API int lshm_stage_planes2d(const float* big, int64_t big_ns, void* planes, int64_t N, int Bc, int h, int w_,
                        lshm_stream_t stream);
LSHM_API int lshm_stage_planes1d(const float* big, int64_t big_ns, void* planes, int64_t N, int Bc, int l, int pad,
                        lshm_stream_t stream);
/* lshm_down2d / lshm_down1d / lshm_wgrad2d / lshm_wgrad1d with the big map given as planes. */
LSHM_API int lshm_down2d_planes(const void* planes, const void* wimg, const float* bias,
                       const float* aux, int64_t aux_ns, float* small_, int64_t small_ns,
                       int64_t N, int A, int Bc, int h, int w_, int epilogue, lshm_stream_t stream);
LSHM_API int lshm_down1d_planes(const void* planes, const void* wimg, const float* bias,
                       const float* aux, int64_t aux_ns, float* small_, int64_t small_ns,
                       int64_t N, int A, int Bc, int l, int epilogue, lshm_stream_t stream);
LSHM_API int lshm_wgrad2d_planes(const float* small_, int64_t small_ns, const void* planes,
                        float* dw, int64_t N, int A, int Bc, int h, int w_, lshm_stream_t stream);
LSHM_API int lshm_wgrad1d_planes(const float* small_, int64_t small_ns, const void* planes,
                        float* dw, int64_t N, int A, int Bc, int l, lshm_stream_t stream);
/* Both gradients of the LAST transposed conv of a 1-D net in one launch, the reconstruction gradient read once:
 * = lshm_wgrad1d_planes(small_, planes -> dw) + lshm_down1d_planes(planes, wimg_down, NULL, aux = small_, ... ->
 * dz, LSHM_EPI_DELU).  A <= 16 small-map channels, Bc in {4, 8}. */
LSHM_API int lshm_tconv_bwd1d_planes(const float* small_, int64_t small_ns, const void* planes, const void* wimg_down,
                            float* dz, int64_t dz_ns, float* dw, int64_t N, int A, int Bc, int l, lshm_stream_t stream);
/* The same for the 2-D net (k4/s2/p1): = lshm_wgrad2d_planes + lshm_down2d_planes(..., aux = small_, LSHM_EPI_DELU).
 * A <= 8 small-map channels, Bc in {4, 8}, w <= 118. */
LSHM_API int lshm_tconv_bwd2d_planes(const float* small_, int64_t small_ns, const void* planes, const void* wimg_down,
                            float* dz, int64_t dz_ns, float* dw, int64_t N, int A, int Bc, int h, int w_, lshm_stream_t stream);
/* The same with the output gradient as an fp32 map `big` (the second-to-last transposed convs, 9..16 -> 8 channels):
 * = lshm_wgrad*d + lshm_down*d(..., aux = small_, LSHM_EPI_DELU), the big map gathered and converted once. */
LSHM_API int lshm_tconv_bwd1d(const float* small_, int64_t small_ns, const float* big, int64_t big_ns, const void* wimg_down,
                     float* dz, int64_t dz_ns, float* dw, int64_t N, int A, int Bc, int l, lshm_stream_t stream);
LSHM_API int lshm_tconv_bwd2d(const float* small_, int64_t small_ns, const float* big, int64_t big_ns, const void* wimg_down,
                     float* dz, int64_t dz_ns, float* dw, int64_t N, int A, int Bc, int h, int w_, lshm_stream_t stream);
/* dst[i] += src[i], i < n (both 16-byte aligned): the parameter gradients of a second micro-batch join the flat
 * gradient buffer before the data-parallel exchange. */
LSHM_API int lshm_vec_add(float* dst, const float* src, int64_t n, lshm_stream_t stream);

/* db[c] = sum over n and positions of g[n,c,:]  (bias gradients). g [N,Cn,len]. */
LSHM_API int lshm_channel_sum(const float* g, int64_t g_ns, float* db, int64_t N, int Cn, int64_t len,
                     lshm_stream_t stream);

/* F.linear (+ELU): y[n,j] = act(b[j] + sum_k x[n*ldx+k] * w[j*K+k])
 * (fcuv1, fc1, fc2in, fc2out, fcuv3, fc3: src/lofar_models.py:80-91). */
LSHM_API int lshm_linear_fwd(const float* x, int64_t ldx, const float* w, const float* b,
                    float* y, int64_t ldy, int64_t N, int K, int J, int epilogue,
                    lshm_stream_t stream);
/* dx[n,k] = (sum_j dz[n,j] w[j,k] + add[n,k]) * (aux ? ELU'(aux[n,k]) : 1)
 * add and aux are nullable and carry their own row strides. */
LSHM_API int lshm_linear_bwd_data(const float* dz, int64_t lddz, const float* w, const float* add,
                         int64_t ldadd, const float* aux, int64_t ldaux, float* dx, int64_t lddx,
                         int64_t N, int K, int J, lshm_stream_t stream);
/* dw[j,k] = sum_n dz[n,j] x[n,k] ; db[j] = sum_n dz[n,j]  (written, not accumulated). */
LSHM_API int lshm_linear_bwd_weight(const float* x, int64_t ldx, const float* dz, int64_t lddz,
                           float* dw, float* db, int64_t N, int K, int J, lshm_stream_t stream);
/* dz = g * ELU'(aux) elementwise over [N,J] with row strides. */
LSHM_API int lshm_delu(const float* g, int64_t ldg, const float* aux, int64_t ldaux, float* dz,
              int64_t lddz, int64_t N, int J, lshm_stream_t stream);

/* ------------------------------------------------------- cascade glue + losses --
 * src/kharmonic_lofar.py:137-147: x11=(x-x1)/2, written as the two 1-D net inputs
 * iyT[n,c,t*P+f] and iyF[n,c,f*P+t]. */
LSHM_API int lshm_residual_split(const float* x, const float* x1, float* iyT, float* iyF,
                        int64_t N, int C, int P, lshm_stream_t stream);
/* lshm_residual_split writing the two 1-D inputs as pad-1 operand planes (Conv1d(k4,s4,p1): window j covers
 * samples [4j-1, 4j+2] of the flattened map) instead of fp32 tensors. */
LSHM_API int lshm_residual_split_planes(const float* x, const float* x1, void* planesT, void* planesF,
                               int64_t N, int C, int P, lshm_stream_t stream);
/* src/kharmonic_lofar.py:150-158.  x,x1 [N,C,P,P]; x2 = netT output viewed [N,C,P,P];
 * x3f = netF output (still transposed: x3[n,c,t,f] = x3f[n,c,f,t]); y1..y3 multipliers
 * in x's flat order.  sums (double[8], caller-zeroed) += { |x1+x2+x3-x|^2, <y1,x-x1>,
 * |x-x1|^2, <y2,x11-x2>, |x11-x2|^2, <y3,x11-x3>, |x11-x3|^2, 0 }.
 * If g1p != NULL also writes the loss gradients scaled by grad_scale (= 1/numel of the
 * GLOBAL batch, so data-parallel shards divide by the same constant as the reference):
 *   g2  = d/dx2,  g3f = d/dx3 stored transposed like x3f,
 *   g1p = direct d/dx1 plus the x11 path of the direct terms (excludes the 1-D nets'
 *         input gradients, added by lshm_cascade_combine).
 * db2 / db3 (both or neither, nullable, float[C], written): per-channel sums of g2 / g3f = the bias
 * gradients of the last transposed conv of the two 1-D nets (saves their lshm_channel_sum passes). */
LSHM_API int lshm_cascade_losses(const float* x, const float* x1, const float* x2, const float* x3f,
                        const float* y1, const float* y2, const float* y3, float rho,
                        int64_t N, int C, int P, float grad_scale, double* sums,
                        float* g1p, float* g2, float* g3f, float* db2, float* db3, lshm_stream_t stream);
/* Same pass with the DEFERRED multiplier update of the previous ADMM iteration folded in
 * (src/kharmonic_lofar.py:187-202 followed by :150-158 on the same minibatch and parameters): when
 * update_y != 0 the kernel first does y_i += rho * r_i with the residuals it computes anyway, stores the
 * new multipliers, and evaluates the loss terms / gradients with them.  update_y == 0 is lshm_cascade_losses.
 * Bit 1 of update_y (values 2, 3): the multipliers are identically ZERO (a new minibatch, :128-130) and are not read;
 * with bit 0 as well the kernel writes y_i = rho * r_i.  (Saves the three memsets and three reads per minibatch.) */
LSHM_API int lshm_cascade_losses_upd(const float* x, const float* x1, const float* x2, const float* x3f,
                            float* y1, float* y2, float* y3, float rho, int update_y,
                            int64_t N, int C, int P, float grad_scale, double* sums,
                            float* g1p, float* g2, float* g3f, float* db2, float* db3, lshm_stream_t stream);
/* lshm_cascade_losses_upd (gradient form) writing g2 / g3 as the pad-0 operand planes of the two 1-D nets' last
 * transposed convs (their dgrad and weight gradient then fetch tiles by TMA); g1p stays fp32. */
LSHM_API int lshm_cascade_losses_planes(const float* x, const float* x1, const float* x2, const float* x3f,
                               float* y1, float* y2, float* y3, float rho, int update_y,
                               int64_t N, int C, int P, float grad_scale, double* sums,
                               float* g1p, void* planes2, void* planes3, float* db2, float* db3, lshm_stream_t stream);
/* gx1 = g1p - 0.5*(gT + transpose(gF)) : total gradient w.r.t. the 2-D net output.
 * db1 (nullable, float[C], written): per-channel sums of gx1 = bias gradient of the 2-D net's last
 * transposed conv. */
LSHM_API int lshm_cascade_combine(const float* g1p, const float* gT, const float* gF, float* gx1,
                         int64_t N, int C, int P, float* db1, lshm_stream_t stream);
/* lshm_cascade_combine writing gx1 as the 2-D operand planes of the 2-D net's last transposed conv. */
LSHM_API int lshm_cascade_combine_planes(const float* g1p, const float* gT, const float* gF, void* planes,
                                int64_t N, int C, int P, float* db1, lshm_stream_t stream);
/* src/kharmonic_lofar.py:200-202: y_i += rho * r_i. */
LSHM_API int lshm_multiplier_update(const float* x, const float* x1, const float* x2, const float* x3f,
                           float rho, float* y1, float* y2, float* y3,
                           int64_t N, int C, int P, lshm_stream_t stream);
/* y_zero != 0: the multipliers start from zero (new minibatch): y_i = rho * r_i is written without reading y_i. */
LSHM_API int lshm_multiplier_update_z(const float* x, const float* x1, const float* x2, const float* x3f,
                             float rho, float* y1, float* y2, float* y3, int y_zero,
                             int64_t N, int C, int P, lshm_stream_t stream);

/* --------------------------------------------------------------- K-harmonic -----
 * Kmeans.forward (src/lofar_models.py:199-209): X [N,L] (row stride ldx), M [K,L].
 * loss_sum (double, caller-zeroed) += sum_i K/(e_i+eps); caller divides by N*K*L.
 * e_out (nullable) [N] receives e_i = sum_k 1/(d_ik^p+eps). */
LSHM_API int lshm_khm_fwd(const float* X, int64_t ldx, const float* M, int64_t N, int K, int L, float p,
                 double* loss_sum, float* e_out, lshm_stream_t stream);
/* Analytic gradient: gX[n,:] (row stride ldg) = or += gscale * dloss_sum/dX,
 * gM [K,L] += gscale * dloss_sum/dM (caller zeroes gM).  accumulate_x: 0 write, 1 add. */
LSHM_API int lshm_khm_bwd(const float* X, int64_t ldx, const float* M, int64_t N, int K, int L, float p,
                 float gscale, float* gX, int64_t ldg, int accumulate_x, float* gM,
                 lshm_stream_t stream);
/* Fused forward+backward in one pass over X (used by the training closure). */
LSHM_API int lshm_khm_fwd_bwd(const float* X, int64_t ldx, const float* M, int64_t N, int K, int L,
                     float p, float gscale, double* loss_sum, float* gX, int64_t ldg,
                     int accumulate_x, float* gM, lshm_stream_t stream);
/* Per-patch nearest centre: ids[n] = argmin_k d_nk (first index on ties). */
LSHM_API int lshm_khm_assign(const float* X, int64_t ldx, const float* M, int64_t N, int K, int L,
                    int32_t* ids, lshm_stream_t stream);
/* src/evaluate_clustering.py:110-119: per group of `group` consecutive rows,
 * dist[g,k] = mean_n d_nk^p ; gid[g] = argmin_k dist[g,k] (first on ties). */
LSHM_API int lshm_khm_group_dist(const float* X, int64_t ldx, const float* M, int64_t N, int K, int L,
                        float p, int group, float* dist, int32_t* gid, lshm_stream_t stream);
/* Centre update sums (Zhang GKHM 7.1-7.5; intent of src/lofar_models.py:231-261):
 * num[K,L] += sum_i Q_ik x_i ; den[K] += sum_i Q_ik (caller zeroes; all-reduce; M=num/den). */
LSHM_API int lshm_khm_center_sums(const float* X, int64_t ldx, const float* M, int64_t N, int K, int L,
                         float p, float* num, float* den, lshm_stream_t stream);
/* M[k,:] = num[k,:] / den[k] */
LSHM_API int lshm_khm_center_apply(const float* num, const float* den, float* M, int K, int L,
                          lshm_stream_t stream);

/* Kmeans.cluster_similarity (src/lofar_models.py:214-229).  loss (double, caller-zeroed)
 * += lscale * similarity ; gM (nullable) += lscale * d similarity/dM.
 * work: float[2*K*K] scratch. */
LSHM_API int lshm_similarity(const float* M, int K, int L, float lscale, double* loss, float* gM,
                    float* work, lshm_stream_t stream);
/* augmented_loss (src/kharmonic_lofar.py:97-110): rows [g*bpb,(g+1)*bpb) form a group.
 * loss += lscale * aug ; gMu (nullable, row stride ldg) += lscale * d aug/dMu. */
LSHM_API int lshm_augment(const float* Mu, int64_t ldx, int64_t N, int L, int bpb, float lscale,
                 double* loss, float* gMu, int64_t ldg, lshm_stream_t stream);
/* RICA penalty term (src/kharmonic_lofar.py:169-171): loss += lscale*sum(log cosh mu)/numel;
 * gmu (nullable) += lscale * tanh(mu)/numel.  mu [N,J] row stride ld. */
LSHM_API int lshm_logcosh(const float* mu, int64_t ld, int64_t N, int J, float lscale, double* loss,
                 float* gmu, int64_t ldg, lshm_stream_t stream);
/* total[0] = sum of the closure's loss terms from the double accumulators:
 * terms[0..6] cascade sums, terms[8] khm sum, terms[9] sim, terms[10] aug, terms[11] rica.
 * Writes out[0]=total, out[1..8] = loss0,loss1,loss2,loss3,kdist,aug,sim,rica (fp32). */
LSHM_API int lshm_closure_total(const double* terms, float rho, double numel, double khm_scale,
                       float* out, lshm_stream_t stream);

/* ------------------------------------------------------------------ optimiser ---
 * torch.optim.Adam semantics (src/kharmonic_lofar.py:92) on one flat buffer. */
LSHM_API int lshm_adam_step(float* p, const float* g, float* m, float* v, int64_t n, float lr,
                   float beta1, float beta2, float eps, int step, lshm_stream_t stream);
/* Same update with the step count kept in device memory: *step_counter is incremented, then used for the
 * bias corrections.  No argument changes from step to step, so the call can be captured in a CUDA graph. */
LSHM_API int lshm_adam_step_dev(float* p, const float* g, float* m, float* v, int64_t n, float lr,
                       float beta1, float beta2, float eps, int32_t* step_counter, lshm_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* LSHM_H_ */
