// Cascade glue, loss terms and their analytic gradients, multiplier update, Adam.
//
// Reference semantics: /root/reference/src/kharmonic_lofar.py:137-172 (closure body),
// :97-110 (augmented_loss), :187-202 (multiplier update), :92 (Adam);
// Kmeans.cluster_similarity /root/reference/src/lofar_models.py:214-229.
//
// The x-sized tensors (x, x1, x2, x3, y1..y3 and three gradients) are each touched exactly
// once per pass, in 32x32 tiles so the transposed frequency-axis tensors go through shared
// memory instead of strided global accesses; the seven loss sums are reduced in registers,
// then per block in double, then one atomic per block.
#include "common.cuh"

namespace lshm {
namespace {

constexpr int TILE = 32, TROWS = 8;

// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TILE * TROWS)
residual_split_kernel(const float* __restrict__ x, const float* __restrict__ x1,
                      float* __restrict__ iyT, float* __restrict__ iyF, int P) {
  __shared__ float tile[TILE][TILE + 1];
  const int64_t plane = (int64_t)blockIdx.z * P * P;
  const int t0 = blockIdx.y * TILE, f0 = blockIdx.x * TILE;
  const int tx = threadIdx.x, ty = threadIdx.y;
#pragma unroll
  for (int i = 0; i < TILE; i += TROWS) {
    const int64_t off = plane + (int64_t)(t0 + ty + i) * P + f0 + tx;
    const float v = 0.5f * (x[off] - x1[off]);
    iyT[off] = v;
    tile[ty + i][tx] = v;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < TILE; i += TROWS)
    iyF[plane + (int64_t)(f0 + ty + i) * P + t0 + tx] = tile[tx][ty + i];
}

// ------------------------------------------------------------------------------------------
constexpr int CASCADE_MAXC = 64;   // channels whose bias-gradient sums can be fused

// 6 blocks per SM (<= 40 registers; 32 spills): at 48 registers the kernel ran at 55 % occupancy and 4.8 TB/s while the
// 30-register multiplier kernel next to it reaches 5.35 TB/s (ncu, profiles/r1_ncu_dram_traffic.json)
// UPD: the deferred multiplier update of the previous ADMM iteration (y_i += rho * r_i,
// src/kharmonic_lofar.py:200-202) is applied on the fly - the residuals r_i are exactly the ones this
// pass computes anyway (same parameters, same minibatch), so the stand-alone 10-pass update kernel and its
// re-read of x, x1, x2, x3 disappear; the loss terms and gradients then use the UPDATED multipliers.
// YZ: the multipliers are identically zero (a new minibatch) and are not read; with UPD the kernel writes y_i = rho r_i.
template <bool GRADS, bool UPD, bool YZ>
__global__ void __launch_bounds__(TILE * TROWS, 6)
cascade_losses_kernel(const float* __restrict__ x, const float* __restrict__ x1,
                      const float* __restrict__ x2, const float* __restrict__ x3f,
                      float* __restrict__ y1, float* __restrict__ y2,
                      float* __restrict__ y3, float rho, float inv_n, int P, int64_t ntiles,
                      double* __restrict__ sums, float* __restrict__ g1p, float* __restrict__ g2,
                      float* __restrict__ g3f, int C, float* __restrict__ db2, float* __restrict__ db3) {
  __shared__ float tile[TILE][TILE + 1];
  __shared__ float gt[TILE][TILE + 1];
  __shared__ double red[32];
  // per-channel sums of g2 / g3f = bias gradients of the last transposed conv of the two 1-D nets
  // (their channel_sum passes would re-read both tensors from HBM)
  __shared__ float cacc[2][CASCADE_MAXC];
  const bool chsum = GRADS && db2 != nullptr;
  if (chsum) {
    for (int c = threadIdx.y * TILE + threadIdx.x; c < 2 * CASCADE_MAXC; c += TILE * TROWS) (&cacc[0][0])[c] = 0.f;
  }
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int tpr = P / TILE, tpp = tpr * tpr;     // tiles per row / per plane
  float s[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  // a block walks a strided list of 32x32 tiles and reduces once at the end
  for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const int64_t plane = (t / tpp) * (int64_t)P * P;
    const int tin = (int)(t % tpp);
    const int t0 = (tin / tpr) * TILE, f0 = (tin % tpr) * TILE;
    float p2 = 0.f, p3 = 0.f;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < TILE; i += TROWS)
      tile[ty + i][tx] = x3f[plane + (int64_t)(f0 + ty + i) * P + t0 + tx];
    __syncthreads();
#pragma unroll
    for (int i = 0; i < TILE; i += TROWS) {
      const int64_t off = plane + (int64_t)(t0 + ty + i) * P + f0 + tx;
      const float xv = x[off], a1 = x1[off], a2 = x2[off], a3 = tile[tx][ty + i];
      float m1 = 0.f, m2 = 0.f, m3 = 0.f;
      if (!YZ) { m1 = y1[off]; m2 = y2[off]; m3 = y3[off]; }
      const float r0 = a1 + a2 + a3 - xv;
      const float r1 = xv - a1;
      const float x11 = 0.5f * r1;
      const float r2 = x11 - a2, r3 = x11 - a3;
      if (UPD) {
        m1 = fmaf(rho, r1, m1); m2 = fmaf(rho, r2, m2); m3 = fmaf(rho, r3, m3);
        y1[off] = m1; y2[off] = m2; y3[off] = m3;
      }
      s[0] = fmaf(r0, r0, s[0]);
      s[1] = fmaf(m1, r1, s[1]); s[2] = fmaf(r1, r1, s[2]);
      s[3] = fmaf(m2, r2, s[3]); s[4] = fmaf(r2, r2, s[4]);
      s[5] = fmaf(m3, r3, s[5]); s[6] = fmaf(r3, r3, s[6]);
      if (GRADS) {
        const float e2 = m2 + rho * r2, e3 = m3 + rho * r3;
        const float v2 = (2.f * r0 - e2) * inv_n, v3 = (2.f * r0 - e3) * inv_n;
        g2[off] = v2;
        gt[ty + i][tx] = v3;
        g1p[off] = (2.f * r0 - m1 - rho * r1 - 0.5f * (e2 + e3)) * inv_n;
        p2 += v2; p3 += v3;
      }
    }
    if (GRADS) {
      __syncthreads();
#pragma unroll
      for (int i = 0; i < TILE; i += TROWS)
        g3f[plane + (int64_t)(f0 + ty + i) * P + t0 + tx] = gt[tx][ty + i];
      if (chsum) {                               // the whole tile belongs to one (sample, channel) plane
        const int c = (int)((t / tpp) % C);
        p2 = warp_sum(p2); p3 = warp_sum(p3);
        if (tx == 0) { atomicAdd(&cacc[0][c], p2); atomicAdd(&cacc[1][c], p3); }
      }
    }
  }
  if (chsum) {
    __syncthreads();
    for (int c = threadIdx.y * TILE + threadIdx.x; c < C; c += TILE * TROWS) {
      atomicAdd(db2 + c, cacc[0][c]); atomicAdd(db3 + c, cacc[1][c]);
    }
  }
  const int tid = ty * TILE + tx;
#pragma unroll
  for (int q = 0; q < 7; ++q) {
    double v = warp_sum((double)s[q]);
    __syncthreads();
    if ((tid & 31) == 0) red[tid >> 5] = v;
    __syncthreads();
    if (tid == 0) {
      double t = 0.0;
      for (int w = 0; w < (TILE * TROWS) / 32; ++w) t += red[w];
      atomicAdd(sums + q, t);
    }
  }
}

// persistent over a strided list of 32x32 tiles; optionally also the per-channel sums of gx1 (the bias
// gradient of the 2-D net's last transposed conv)
__global__ void __launch_bounds__(TILE * TROWS)
cascade_combine_kernel(const float* __restrict__ g1p, const float* __restrict__ gT,
                       const float* __restrict__ gF, float* __restrict__ gx1, int P, int64_t ntiles, int C,
                       float* __restrict__ db1) {
  __shared__ float tile[TILE][TILE + 1];
  __shared__ float cacc[CASCADE_MAXC];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int tpr = P / TILE, tpp = tpr * tpr;
  if (db1 != nullptr)
    for (int c = ty * TILE + tx; c < CASCADE_MAXC; c += TILE * TROWS) cacc[c] = 0.f;
  for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const int64_t plane = (t / tpp) * (int64_t)P * P;
    const int tin = (int)(t % tpp);
    const int t0 = (tin / tpr) * TILE, f0 = (tin % tpr) * TILE;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < TILE; i += TROWS)
      tile[ty + i][tx] = gF[plane + (int64_t)(f0 + ty + i) * P + t0 + tx];
    __syncthreads();
    float p1 = 0.f;
#pragma unroll
    for (int i = 0; i < TILE; i += TROWS) {
      const int64_t off = plane + (int64_t)(t0 + ty + i) * P + f0 + tx;
      const float v = g1p[off] - 0.5f * (gT[off] + tile[tx][ty + i]);
      gx1[off] = v;
      p1 += v;
    }
    if (db1 != nullptr) {
      p1 = warp_sum(p1);
      if (tx == 0) atomicAdd(&cacc[(int)((t / tpp) % C)], p1);
    }
  }
  if (db1 != nullptr) {
    __syncthreads();
    for (int c = ty * TILE + tx; c < C; c += TILE * TROWS) atomicAdd(db1 + c, cacc[c]);
  }
}

template <bool YZ>
__global__ void __launch_bounds__(TILE * TROWS)
multiplier_update_kernel(const float* __restrict__ x, const float* __restrict__ x1,
                         const float* __restrict__ x2, const float* __restrict__ x3f, float rho,
                         float* __restrict__ y1, float* __restrict__ y2, float* __restrict__ y3, int P) {
  __shared__ float tile[TILE][TILE + 1];
  const int64_t plane = (int64_t)blockIdx.z * P * P;
  const int t0 = blockIdx.y * TILE, f0 = blockIdx.x * TILE;
  const int tx = threadIdx.x, ty = threadIdx.y;
#pragma unroll
  for (int i = 0; i < TILE; i += TROWS)
    tile[ty + i][tx] = x3f[plane + (int64_t)(f0 + ty + i) * P + t0 + tx];
  __syncthreads();
#pragma unroll
  for (int i = 0; i < TILE; i += TROWS) {
    const int64_t off = plane + (int64_t)(t0 + ty + i) * P + f0 + tx;
    const float r1 = x[off] - x1[off];
    const float x11 = 0.5f * r1;
    if (YZ) {       // multipliers start from zero (new minibatch): written, not read
      y1[off] = rho * r1;
      y2[off] = rho * (x11 - x2[off]);
      y3[off] = rho * (x11 - tile[tx][ty + i]);
    } else {
      y1[off] += rho * r1;
      y2[off] += rho * (x11 - x2[off]);
      y3[off] += rho * (x11 - tile[tx][ty + i]);
    }
  }
}

// ------------------------------------------------------------------------------------------
// cluster similarity
__global__ void gram_kernel(const float* __restrict__ M, int K, int L, float* __restrict__ G) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= K * K) return;
  const int i = idx / K, j = idx - i * K;
  float s = 0.f;
  for (int l = 0; l < L; ++l) s = fmaf(M[(int64_t)i * L + l], M[(int64_t)j * L + l], s);
  G[idx] = s;
}

__global__ void __launch_bounds__(128)
similarity_kernel(const float* __restrict__ M, const float* __restrict__ G, int K, int L,
                  float lscale, double* __restrict__ loss, float* __restrict__ gM,
                  float* __restrict__ alpha_ws) {
  constexpr float EPS = 1e-9f;
  __shared__ float red[32];
  __shared__ float sh_beta, sh_numer;
  const int r = blockIdx.x;
  float* alpha = alpha_ws + (int64_t)r * K;   // per-centre row of coefficients (global scratch)
  const float Grr = G[(int64_t)r * K + r];
  const float nr = sqrtf(Grr);
  const float Dr = expf(Grr / (nr * nr + EPS));
  const float inv_r = 1.f / (Dr + EPS);
  float numer = 0.f, beta = 0.f;
  for (int j = threadIdx.x; j < K; j += blockDim.x) {
    float a = 0.f;
    if (j != r) {
      const float Gjj = G[(int64_t)j * K + j], Grj = G[(int64_t)r * K + j];
      const float nj = sqrtf(Gjj);
      const float den = nr * nj + EPS;
      const float S = expf(Grj / den);
      const float Dj = expf(Gjj / (nj * nj + EPS));
      const float c = S * (inv_r + 1.f / (Dj + EPS));
      numer += S;
      a = c / den;
      beta -= c * Grj * nj / (nr * den * den);
    }
    alpha[j] = a;
  }
  numer = block_sum<float>(numer, red);
  if (threadIdx.x == 0) sh_numer = numer;
  beta = block_sum<float>(beta, red);
  if (threadIdx.x == 0) sh_beta = beta;
  __syncthreads();
  numer = sh_numer;
  const float n2e = nr * nr + EPS;
  beta = sh_beta - numer * Dr * inv_r * inv_r * (2.f * EPS / (n2e * n2e));
  const float scale = lscale / ((float)K * (float)L);
  if (threadIdx.x == 0) atomicAdd(loss, (double)(scale * numer * inv_r));
  if (gM != nullptr) {
    for (int l = threadIdx.x; l < L; l += blockDim.x) {
      float g = beta * M[(int64_t)r * L + l];
      for (int j = 0; j < K; ++j) g = fmaf(alpha[j], M[(int64_t)j * L + l], g);
      gM[(int64_t)r * L + l] += scale * g;
    }
  }
}

// ------------------------------------------------------------------------------------------
// augmentation loss: one block per baseline group
__global__ void __launch_bounds__(128)
augment_kernel(const float* __restrict__ Mu, int64_t ldx, int L, int bpb, float lscale,
               double* __restrict__ loss, float* __restrict__ gMu, int64_t ldg) {
  extern __shared__ float sm[];
  float* zh = sm;                       // [bpb][L] normalised rows
  float* nrm = zh + (size_t)bpb * L;    // [bpb]
  float* E = nrm + bpb;                 // [bpb][bpb]
  __shared__ float red[32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int64_t row0 = (int64_t)blockIdx.x * bpb;
  for (int i = wid; i < bpb; i += nw) {
    const float* z = Mu + (row0 + i) * ldx;
    float s = 0.f;
    for (int l = lane; l < L; l += 32) s = fmaf(z[l], z[l], s);
    s = sqrtf(warp_sum(s));
    const float inv = 1.f / (s + 1e-6f);
    for (int l = lane; l < L; l += 32) zh[(size_t)i * L + l] = z[l] * inv;
    if (lane == 0) nrm[i] = s;
  }
  __syncthreads();
  float lsum = 0.f;
  for (int pair = wid; pair < bpb * bpb; pair += nw) {
    const int i = pair / bpb, j = pair - i * bpb;
    if (j <= i) continue;
    float d = 0.f;
    for (int l = lane; l < L; l += 32) d = fmaf(zh[(size_t)i * L + l], zh[(size_t)j * L + l], d);
    d = warp_sum(d);
    const float e = expf(-d);
    if (lane == 0) { E[i * bpb + j] = e; E[j * bpb + i] = e; lsum += e; }
  }
  lsum = block_sum<float>(lsum, red);
  if (threadIdx.x == 0) atomicAdd(loss, (double)(lscale * lsum));
  if (gMu == nullptr) return;
  __syncthreads();
  for (int i = wid; i < bpb; i += nw) {
    const float n = nrm[i], nd = n + 1e-6f;
    // g wrt normalised row, then through z/(|z|+delta)
    float dot = 0.f;
    for (int l = lane; l < L; l += 32) {
      float g = 0.f;
      for (int j = 0; j < bpb; ++j)
        if (j != i) g = fmaf(-E[i * bpb + j], zh[(size_t)j * L + l], g);
      dot = fmaf(zh[(size_t)i * L + l] * nd, g, dot);
    }
    dot = warp_sum(dot);
    const float c2 = n > 0.f ? dot / (n * nd * nd) : 0.f;
    for (int l = lane; l < L; l += 32) {
      float g = 0.f;
      for (int j = 0; j < bpb; ++j)
        if (j != i) g = fmaf(-E[i * bpb + j], zh[(size_t)j * L + l], g);
      const float zi = zh[(size_t)i * L + l] * nd;
      gMu[(row0 + i) * ldg + l] += lscale * (g / nd - zi * c2);
    }
  }
}

// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
logcosh_kernel(const float* __restrict__ mu, int64_t ld, int64_t N, int J, float lscale,
               double* __restrict__ loss, float* __restrict__ gmu, int64_t ldg) {
  __shared__ double red[32];
  const int64_t total = N * J;
  const float inv = lscale / (float)total;
  double s = 0.0;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t n = idx / J;
    const int j = (int)(idx - n * J);
    const float v = mu[n * ld + j];
    s += (double)logf(coshf(v));
    if (gmu != nullptr) gmu[n * ldg + j] += inv * tanhf(v);
  }
  s = block_sum<double>(s, red);
  if (threadIdx.x == 0) atomicAdd(loss, s * (double)inv);
}

__global__ void closure_total_kernel(const double* __restrict__ t, float rho, double numel,
                                     double khm_scale, float* __restrict__ out) {
  const double l0 = t[0] / numel;
  const double l1 = (t[1] + 0.5 * rho * t[2]) / numel;
  const double l2 = (t[3] + 0.5 * rho * t[4]) / numel;
  const double l3 = (t[5] + 0.5 * rho * t[6]) / numel;
  const double kd = t[8] * khm_scale, sim = t[9], aug = t[10], rica = t[11];
  out[0] = (float)(l0 + l1 + l2 + l3 + kd + aug + sim + rica);
  out[1] = (float)l0; out[2] = (float)l1; out[3] = (float)l2; out[4] = (float)l3;
  out[5] = (float)kd; out[6] = (float)aug; out[7] = (float)sim; out[8] = (float)rica;
}

__global__ void counter_inc_kernel(int32_t* c) { *c += 1; }

// Adam with the step count in device memory (incremented by counter_inc_kernel just before): nothing
// step-dependent is a kernel argument, so the launch can be replayed from a CUDA graph.
__global__ void __launch_bounds__(256)
adam_dev_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                float* __restrict__ v, int64_t n, float lr, float b1, float b2, float eps,
                const int32_t* __restrict__ step) {
  const double t = (double)*step;
  const float bc1 = (float)(1.0 - pow((double)b1, t));
  const float bc2_sqrt = sqrtf((float)(1.0 - pow((double)b2, t)));
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const float gi = g[i];
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi; v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] -= (lr / bc1) * (mi / denom);
  }
}

__global__ void __launch_bounds__(256)
adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
            float* __restrict__ v, int64_t n, float lr, float b1, float b2, float eps,
            float bc1, float bc2_sqrt) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const float gi = g[i];
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi; v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] -= (lr / bc1) * (mi / denom);
  }
}

int check_cascade(const char* name, int64_t N, int C, int P) {
  LSHM_REQUIRE(N >= 0 && C > 0 && P > 0 && P % TILE == 0, "%s: need P%%32==0 (N=%lld C=%d P=%d)", name, (long long)N, C, P);
  return LSHM_OK;
}

}  // namespace
}  // namespace lshm

using namespace lshm;

extern "C" {

int lshm_residual_split(const float* x, const float* x1, float* iyT, float* iyF,
                        int64_t N, int C, int P, lshm_stream_t stream) {
  LSHM_REQUIRE(x && x1 && iyT && iyF, "lshm_residual_split: null pointer");
  if (int rc = check_cascade("lshm_residual_split", N, C, P)) return rc;
  if (N == 0) return LSHM_OK;
  // blockIdx.z is limited to 65535: split planes over several launches
  const int64_t planes = N * C;
  for (int64_t z0 = 0; z0 < planes; z0 += 65535) {
    const int64_t nz = planes - z0 < 65535 ? planes - z0 : 65535;
    const int64_t o = z0 * P * P;
    dim3 grid(P / TILE, P / TILE, (unsigned)nz), block(TILE, TROWS);
    residual_split_kernel<<<grid, block, 0, as_stream(stream)>>>(x + o, x1 + o, iyT + o, iyF + o, P);
  }
  LSHM_CHECK_LAUNCH("lshm_residual_split");
  return LSHM_OK;
}

int lshm_cascade_losses(const float* x, const float* x1, const float* x2, const float* x3f,
                        const float* y1, const float* y2, const float* y3, float rho,
                        int64_t N, int C, int P, float grad_scale, double* sums,
                        float* g1p, float* g2, float* g3f, float* db2, float* db3, lshm_stream_t stream) {
  return lshm_cascade_losses_upd(x, x1, x2, x3f, const_cast<float*>(y1), const_cast<float*>(y2), const_cast<float*>(y3),
                                 rho, 0, N, C, P, grad_scale, sums, g1p, g2, g3f, db2, db3, stream);
}

int lshm_cascade_losses_upd(const float* x, const float* x1, const float* x2, const float* x3f,
                            float* y1, float* y2, float* y3, float rho, int update_y,
                            int64_t N, int C, int P, float grad_scale, double* sums,
                            float* g1p, float* g2, float* g3f, float* db2, float* db3, lshm_stream_t stream) {
  LSHM_REQUIRE(x && x1 && x2 && x3f && y1 && y2 && y3 && sums, "lshm_cascade_losses: null pointer");
  LSHM_REQUIRE((db2 == nullptr) == (db3 == nullptr) && (db2 == nullptr || (g1p != nullptr && C <= CASCADE_MAXC)),
               "lshm_cascade_losses: db2/db3 come together, need the gradient outputs and C <= %d", CASCADE_MAXC);
  LSHM_REQUIRE((g1p == nullptr) == (g2 == nullptr) && (g1p == nullptr) == (g3f == nullptr),
               "lshm_cascade_losses: give all three gradient outputs or none");
  if (int rc = check_cascade("lshm_cascade_losses", N, C, P)) return rc;
  if (N == 0) return LSHM_OK;
  const float inv_n = grad_scale;
  const int64_t ntiles = N * C * (int64_t)(P / TILE) * (P / TILE);
  const int64_t blocks = std::min<int64_t>(ntiles, (int64_t)sm_count() * 16);
  dim3 block(TILE, TROWS);
  if (db2) {
    LSHM_CUDA(cudaMemsetAsync(db2, 0, sizeof(float) * C, as_stream(stream)), "lshm_cascade_losses");
    LSHM_CUDA(cudaMemsetAsync(db3, 0, sizeof(float) * C, as_stream(stream)), "lshm_cascade_losses");
  }
#define LSHM_CL(G, U, Z) cascade_losses_kernel<G, U, Z><<<(unsigned)blocks, block, 0, as_stream(stream)>>>( \
      x, x1, x2, x3f, y1, y2, y3, rho, inv_n, P, ntiles, sums, g1p, g2, g3f, C, db2, db3)
  const int mode = update_y & 3;     // bit 0: apply the deferred update; bit 1: the multipliers are zero (not read)
  if (g1p) {
    if (mode == 0) LSHM_CL(true, false, false); else if (mode == 1) LSHM_CL(true, true, false);
    else if (mode == 2) LSHM_CL(true, false, true); else LSHM_CL(true, true, true);
  } else {
    if (mode == 0) LSHM_CL(false, false, false); else if (mode == 1) LSHM_CL(false, true, false);
    else if (mode == 2) LSHM_CL(false, false, true); else LSHM_CL(false, true, true);
  }
#undef LSHM_CL
  LSHM_CHECK_LAUNCH("lshm_cascade_losses");
  return LSHM_OK;
}

int lshm_cascade_combine(const float* g1p, const float* gT, const float* gF, float* gx1,
                         int64_t N, int C, int P, float* db1, lshm_stream_t stream) {
  LSHM_REQUIRE(g1p && gT && gF && gx1, "lshm_cascade_combine: null pointer");
  LSHM_REQUIRE(db1 == nullptr || C <= CASCADE_MAXC, "lshm_cascade_combine: db1 needs C <= %d", CASCADE_MAXC);
  if (int rc = check_cascade("lshm_cascade_combine", N, C, P)) return rc;
  if (N == 0) return LSHM_OK;
  if (db1) LSHM_CUDA(cudaMemsetAsync(db1, 0, sizeof(float) * C, as_stream(stream)), "lshm_cascade_combine");
  const int64_t ntiles = N * C * (int64_t)(P / TILE) * (P / TILE);
  const int64_t blocks = std::min<int64_t>(ntiles, (int64_t)sm_count() * 16);
  dim3 block(TILE, TROWS);
  cascade_combine_kernel<<<(unsigned)blocks, block, 0, as_stream(stream)>>>(g1p, gT, gF, gx1, P, ntiles, C, db1);
  LSHM_CHECK_LAUNCH("lshm_cascade_combine");
  return LSHM_OK;
}

int lshm_multiplier_update(const float* x, const float* x1, const float* x2, const float* x3f,
                           float rho, float* y1, float* y2, float* y3,
                           int64_t N, int C, int P, lshm_stream_t stream) {
  return lshm_multiplier_update_z(x, x1, x2, x3f, rho, y1, y2, y3, 0, N, C, P, stream);
}

int lshm_multiplier_update_z(const float* x, const float* x1, const float* x2, const float* x3f,
                             float rho, float* y1, float* y2, float* y3, int y_zero,
                             int64_t N, int C, int P, lshm_stream_t stream) {
  LSHM_REQUIRE(x && x1 && x2 && x3f && y1 && y2 && y3, "lshm_multiplier_update: null pointer");
  if (int rc = check_cascade("lshm_multiplier_update", N, C, P)) return rc;
  if (N == 0) return LSHM_OK;
  const int64_t planes = N * C;
  for (int64_t z0 = 0; z0 < planes; z0 += 65535) {
    const int64_t nz = planes - z0 < 65535 ? planes - z0 : 65535;
    const int64_t o = z0 * P * P;
    dim3 grid(P / TILE, P / TILE, (unsigned)nz), block(TILE, TROWS);
    if (y_zero) multiplier_update_kernel<true><<<grid, block, 0, as_stream(stream)>>>(x + o, x1 + o, x2 + o, x3f + o, rho, y1 + o, y2 + o, y3 + o, P);
    else multiplier_update_kernel<false><<<grid, block, 0, as_stream(stream)>>>(x + o, x1 + o, x2 + o, x3f + o, rho, y1 + o, y2 + o, y3 + o, P);
  }
  LSHM_CHECK_LAUNCH("lshm_multiplier_update");
  return LSHM_OK;
}

int lshm_similarity(const float* M, int K, int L, float lscale, double* loss, float* gM,
                    float* work, lshm_stream_t stream) {
  LSHM_REQUIRE(M && loss && work && K > 0 && L > 0, "lshm_similarity: bad arguments");
  LSHM_REQUIRE(K <= 16384, "lshm_similarity: K too large");
  cudaStream_t st = as_stream(stream);
  gram_kernel<<<(unsigned)ceil_div((int64_t)K * K, 256), 256, 0, st>>>(M, K, L, work);
  // the second K*K block of `work` holds the per-centre coefficient rows
  similarity_kernel<<<K, 128, 0, st>>>(M, work, K, L, lscale, loss, gM, work + (size_t)K * K);
  LSHM_CHECK_LAUNCH("lshm_similarity");
  return LSHM_OK;
}

int lshm_augment(const float* Mu, int64_t ldx, int64_t N, int L, int bpb, float lscale,
                 double* loss, float* gMu, int64_t ldg, lshm_stream_t stream) {
  LSHM_REQUIRE(Mu && loss && L > 0 && bpb > 0 && N >= 0, "lshm_augment: bad arguments");
  LSHM_REQUIRE(N % bpb == 0, "lshm_augment: N=%lld is not a multiple of bpb=%d", (long long)N, bpb);
  if (N == 0) return LSHM_OK;
  const size_t smem = ((size_t)bpb * L + bpb + (size_t)bpb * bpb) * sizeof(float);
  LSHM_REQUIRE(smem <= 200 * 1024, "lshm_augment: group too large for shared memory (bpb=%d L=%d)", bpb, L);
  if (smem > 48 * 1024)
    LSHM_CUDA(cudaFuncSetAttribute(augment_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "lshm_augment");
  augment_kernel<<<(unsigned)(N / bpb), 128, smem, as_stream(stream)>>>(Mu, ldx, L, bpb, lscale, loss, gMu, ldg);
  LSHM_CHECK_LAUNCH("lshm_augment");
  return LSHM_OK;
}

int lshm_logcosh(const float* mu, int64_t ld, int64_t N, int J, float lscale, double* loss,
                 float* gmu, int64_t ldg, lshm_stream_t stream) {
  LSHM_REQUIRE(mu && loss && N >= 0 && J > 0, "lshm_logcosh: bad arguments");
  if (N == 0) return LSHM_OK;
  const int64_t blocks = std::min<int64_t>(ceil_div(N * J, 256), (int64_t)sm_count() * 8);
  logcosh_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(mu, ld, N, J, lscale, loss, gmu, ldg);
  LSHM_CHECK_LAUNCH("lshm_logcosh");
  return LSHM_OK;
}

int lshm_closure_total(const double* terms, float rho, double numel, double khm_scale,
                       float* out, lshm_stream_t stream) {
  LSHM_REQUIRE(terms && out && numel > 0, "lshm_closure_total: bad arguments");
  closure_total_kernel<<<1, 1, 0, as_stream(stream)>>>(terms, rho, numel, khm_scale, out);
  LSHM_CHECK_LAUNCH("lshm_closure_total");
  return LSHM_OK;
}

int lshm_adam_step_dev(float* p, const float* g, float* m, float* v, int64_t n, float lr,
                       float beta1, float beta2, float eps, int32_t* step_counter, lshm_stream_t stream) {
  LSHM_REQUIRE(p && g && m && v && step_counter && n >= 0, "lshm_adam_step_dev: bad arguments");
  cudaStream_t st = as_stream(stream);
  counter_inc_kernel<<<1, 1, 0, st>>>(step_counter);
  LSHM_CHECK_LAUNCH("lshm_adam_step_dev");
  if (n == 0) return LSHM_OK;
  const int64_t blocks = std::min<int64_t>(ceil_div(n, 256), (int64_t)sm_count() * 8);
  adam_dev_kernel<<<(unsigned)blocks, 256, 0, st>>>(p, g, m, v, n, lr, beta1, beta2, eps, step_counter);
  LSHM_CHECK_LAUNCH("lshm_adam_step_dev");
  return LSHM_OK;
}

int lshm_adam_step(float* p, const float* g, float* m, float* v, int64_t n, float lr,
                   float beta1, float beta2, float eps, int step, lshm_stream_t stream) {
  LSHM_REQUIRE(p && g && m && v && n >= 0 && step >= 1, "lshm_adam_step: bad arguments");
  if (n == 0) return LSHM_OK;
  const float bc1 = 1.f - powf(beta1, (float)step);
  const float bc2s = sqrtf(1.f - powf(beta2, (float)step));
  const int64_t blocks = std::min<int64_t>(ceil_div(n, 256), (int64_t)sm_count() * 8);
  adam_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(p, g, m, v, n, lr, beta1, beta2, eps, bc1, bc2s);
  LSHM_CHECK_LAUNCH("lshm_adam_step");
  return LSHM_OK;
}

}  // extern "C"
