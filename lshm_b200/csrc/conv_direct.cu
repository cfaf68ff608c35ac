// Direct (CUDA-core, fp32) kernels for the k4/s2 2-D and k4/s4 1-D convolutions and their
// transposes: "down" (Conv fwd / ConvTranspose dgrad), "up" (ConvTranspose fwd / Conv dgrad)
// and "wgrad".  Geometry and naming: see include/lshm.h.
//
// Reference semantics: F.conv2d/conv_transpose2d (k4,s2,p1) at
// /root/reference/src/lofar_models.py:73-78,93-98 and F.conv1d (k4,s4,p1) /
// conv_transpose1d (k4,s4,p0) at :158-163,:178-183; F.elu fused into the epilogue.
//
// Each thread owns one output pixel and a small register tile of output channels; the weight
// slice of the block is staged once in shared memory, transposed so the channel tile of one tap
// is a broadcast float4.
#include "common.cuh"

namespace lshm {
namespace {

constexpr int CONV_THREADS = 128;

template <int T>
__device__ __forceinline__ void fma_tile(float (&acc)[T], float v, const float* wrow) {
#pragma unroll
  for (int i = 0; i < T; i += 4) {
    const float4 wv = *reinterpret_cast<const float4*>(wrow + i);
    acc[i + 0] = fmaf(v, wv.x, acc[i + 0]);
    acc[i + 1] = fmaf(v, wv.y, acc[i + 1]);
    acc[i + 2] = fmaf(v, wv.z, acc[i + 2]);
    acc[i + 3] = fmaf(v, wv.w, acc[i + 3]);
  }
}

__device__ __forceinline__ float apply_epi(float acc, int epi, const float* aux, int64_t off) {
  if (epi == LSHM_EPI_ELU) return elu_f(acc);
  if (epi == LSHM_EPI_DELU) return acc * delu_from_out(aux[off]);
  return acc;
}

// ------------------------------------------------------------------ 2-D down -------------
template <int AT>
__global__ void __launch_bounds__(CONV_THREADS)
down2d_kernel(const float* __restrict__ big, int64_t big_ns, const float* __restrict__ w,
              const float* __restrict__ bias, const float* __restrict__ aux, int64_t aux_ns,
              float* __restrict__ small_, int64_t small_ns, int64_t N, int A, int Bc, int h, int w_, int epi) {
  extern __shared__ __align__(16) float ws[];  // [Bc][16][AT]
  const int a0 = blockIdx.y * AT;
  for (int idx = threadIdx.x; idx < AT * Bc * 16; idx += blockDim.x) {
    const int ai = idx / (Bc * 16), rem = idx - ai * (Bc * 16);
    const int a = a0 + ai;
    ws[rem * AT + ai] = a < A ? w[(int64_t)a * Bc * 16 + rem] : 0.f;
  }
  __syncthreads();
  const int64_t hw = (int64_t)h * w_;
  const int64_t pix = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= N * hw) return;
  const int64_t n = pix / hw;
  const int r = (int)(pix - n * hw);
  const int oy = r / w_, ox = r - oy * w_;
  const int H = 2 * h, W = 2 * w_;
  float acc[AT];
#pragma unroll
  for (int i = 0; i < AT; ++i) acc[i] = (bias != nullptr && a0 + i < A) ? bias[a0 + i] : 0.f;
  const float* bp = big + n * big_ns;
  const int iy0 = 2 * oy - 1, ix0 = 2 * ox - 1;
  for (int b = 0; b < Bc; ++b) {
    const float* plane = bp + (int64_t)b * H * W;
#pragma unroll
    for (int ky = 0; ky < 4; ++ky) {
      const int iy = iy0 + ky;
      const bool yin = iy >= 0 && iy < H;
#pragma unroll
      for (int kx = 0; kx < 4; ++kx) {
        const int ix = ix0 + kx;
        const float v = (yin && ix >= 0 && ix < W) ? __ldg(plane + (int64_t)iy * W + ix) : 0.f;
        fma_tile<AT>(acc, v, ws + ((b * 16 + ky * 4 + kx) * AT));
      }
    }
  }
#pragma unroll
  for (int i = 0; i < AT; ++i) {
    const int a = a0 + i;
    if (a < A) {
      const int64_t off = (int64_t)a * hw + r;
      small_[n * small_ns + off] = apply_epi(acc[i], epi, aux + n * aux_ns, off);
    }
  }
}

// ------------------------------------------------------------------ 2-D up ---------------
template <int BT>
__global__ void __launch_bounds__(CONV_THREADS)
up2d_kernel(const float* __restrict__ small_, int64_t small_ns, const float* __restrict__ w,
            const float* __restrict__ bias, const float* __restrict__ aux, int64_t aux_ns,
            float* __restrict__ big, int64_t big_ns, int64_t N, int A, int Bc, int h, int w_, int epi) {
  extern __shared__ __align__(16) float ws[];  // [A][16][BT]
  const int b0 = blockIdx.y * BT;
  for (int idx = threadIdx.x; idx < A * BT * 16; idx += blockDim.x) {
    const int a = idx / (BT * 16), rem = idx - a * (BT * 16);
    const int bi = rem >> 4, tap = rem & 15;
    const int b = b0 + bi;
    ws[(a * 16 + tap) * BT + bi] = b < Bc ? w[((int64_t)a * Bc + b) * 16 + tap] : 0.f;
  }
  __syncthreads();
  const int H = 2 * h, W = 2 * w_;
  const int64_t HW = (int64_t)H * W, hw = (int64_t)h * w_;
  const int64_t pix = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= N * HW) return;
  const int64_t n = pix / HW;
  const int r = (int)(pix - n * HW);
  const int y = r / W, x = r - y * W;
  const int py = (y + 1) & 1, px = (x + 1) & 1;
  float acc[BT];
#pragma unroll
  for (int i = 0; i < BT; ++i) acc[i] = (bias != nullptr && b0 + i < Bc) ? bias[b0 + i] : 0.f;
  const float* sp = small_ + n * small_ns;
  // the two contributing rows / columns of the small map
  int oys[2], kys[2], oxs[2], kxs[2];
  bool yok[2], xok[2];
#pragma unroll
  for (int d = 0; d < 2; ++d) {
    kys[d] = py + 2 * d; oys[d] = (y + 1 - kys[d]) / 2; yok[d] = (y + 1 - kys[d]) >= 0 && oys[d] < h;
    kxs[d] = px + 2 * d; oxs[d] = (x + 1 - kxs[d]) / 2; xok[d] = (x + 1 - kxs[d]) >= 0 && oxs[d] < w_;
  }
  for (int a = 0; a < A; ++a) {
    const float* plane = sp + (int64_t)a * hw;
#pragma unroll
    for (int dy = 0; dy < 2; ++dy) {
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {
        const float v = (yok[dy] && xok[dx]) ? __ldg(plane + oys[dy] * w_ + oxs[dx]) : 0.f;
        fma_tile<BT>(acc, v, ws + ((a * 16 + kys[dy] * 4 + kxs[dx]) * BT));
      }
    }
  }
#pragma unroll
  for (int i = 0; i < BT; ++i) {
    const int b = b0 + i;
    if (b < Bc) {
      const int64_t off = (int64_t)b * HW + r;
      big[n * big_ns + off] = apply_epi(acc[i], epi, aux + n * aux_ns, off);
    }
  }
}

// ------------------------------------------------------------------ 2-D wgrad ------------
// dW[a,b,ky,kx] = sum_{n,oy,ox} S[n,a,oy,ox] * B[n,b,2oy-1+ky,2ox-1+kx]
// block: one (a-tile, b) pair and a chunk of pixels; thread keeps AT x 16 partial sums.
constexpr int WG_THREADS = 256;
template <int AT>
__global__ void __launch_bounds__(WG_THREADS)
wgrad2d_kernel(const float* __restrict__ small_, int64_t small_ns, const float* __restrict__ big,
               int64_t big_ns, float* __restrict__ dw, int64_t N, int A, int Bc, int h, int w_,
               int64_t chunk) {
  __shared__ float red[WG_THREADS / 32][AT * 16];
  const int combo = blockIdx.y;
  const int b = combo % Bc, a0 = (combo / Bc) * AT;
  const int H = 2 * h, W = 2 * w_;
  const int64_t hw = (int64_t)h * w_, total = N * hw;
  const int64_t start = (int64_t)blockIdx.x * chunk;
  const int64_t stop = min(start + chunk, total);
  float acc[AT][16];
#pragma unroll
  for (int i = 0; i < AT; ++i)
#pragma unroll
    for (int t = 0; t < 16; ++t) acc[i][t] = 0.f;
  for (int64_t pix = start + threadIdx.x; pix < stop; pix += WG_THREADS) {
    const int64_t n = pix / hw;
    const int r = (int)(pix - n * hw);
    const int oy = r / w_, ox = r - oy * w_;
    const float* plane = big + n * big_ns + (int64_t)b * H * W;
    float win[16];
#pragma unroll
    for (int ky = 0; ky < 4; ++ky) {
      const int iy = 2 * oy - 1 + ky;
      const bool yin = iy >= 0 && iy < H;
#pragma unroll
      for (int kx = 0; kx < 4; ++kx) {
        const int ix = 2 * ox - 1 + kx;
        win[ky * 4 + kx] = (yin && ix >= 0 && ix < W) ? __ldg(plane + (int64_t)iy * W + ix) : 0.f;
      }
    }
    const float* sp = small_ + n * small_ns + r;
#pragma unroll
    for (int i = 0; i < AT; ++i) {
      const float s = (a0 + i < A) ? __ldg(sp + (int64_t)(a0 + i) * hw) : 0.f;
#pragma unroll
      for (int t = 0; t < 16; ++t) acc[i][t] = fmaf(s, win[t], acc[i][t]);
    }
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < AT; ++i)
#pragma unroll
    for (int t = 0; t < 16; ++t) {
      const float v = warp_sum(acc[i][t]);
      if (lane == 0) red[wid][i * 16 + t] = v;
    }
  __syncthreads();
  if (threadIdx.x < AT * 16) {
    float v = 0.f;
#pragma unroll
    for (int q = 0; q < WG_THREADS / 32; ++q) v += red[q][threadIdx.x];
    const int i = threadIdx.x >> 4, t = threadIdx.x & 15;
    if (a0 + i < A) atomicAdd(dw + ((int64_t)(a0 + i) * Bc + b) * 16 + t, v);
  }
}

// ------------------------------------------------------------------ 1-D down -------------
template <int AT>
__global__ void __launch_bounds__(CONV_THREADS)
down1d_kernel(const float* __restrict__ big, int64_t big_ns, const float* __restrict__ w,
              const float* __restrict__ bias, const float* __restrict__ aux, int64_t aux_ns,
              float* __restrict__ small_, int64_t small_ns, int64_t N, int A, int Bc, int l, int pad,
              int epi, int vec_ok) {
  extern __shared__ __align__(16) float ws[];  // [Bc][4][AT]
  const int a0 = blockIdx.y * AT;
  for (int idx = threadIdx.x; idx < AT * Bc * 4; idx += blockDim.x) {
    const int ai = idx / (Bc * 4), rem = idx - ai * (Bc * 4);
    const int a = a0 + ai;
    ws[rem * AT + ai] = a < A ? w[(int64_t)a * Bc * 4 + rem] : 0.f;
  }
  __syncthreads();
  const int64_t pix = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= N * l) return;
  const int64_t n = pix / l;
  const int j = (int)(pix - n * l);
  const int64_t Lb = 4 * (int64_t)l;
  float acc[AT];
#pragma unroll
  for (int i = 0; i < AT; ++i) acc[i] = (bias != nullptr && a0 + i < A) ? bias[a0 + i] : 0.f;
  const float* bp = big + n * big_ns + 4 * (int64_t)j;
  for (int b = 0; b < Bc; ++b) {
    const float* row = bp + (int64_t)b * Lb;
    float v0, v1, v2, v3;
    if (pad == 0) {
      if (vec_ok) {
        const float4 q = __ldg(reinterpret_cast<const float4*>(row));
        v0 = q.x; v1 = q.y; v2 = q.z; v3 = q.w;
      } else {
        v0 = __ldg(row); v1 = __ldg(row + 1); v2 = __ldg(row + 2); v3 = __ldg(row + 3);
      }
    } else {  // window [4j-1, 4j+2]; position -1 is the zero pad
      v0 = j > 0 ? __ldg(row - 1) : 0.f;
      v1 = __ldg(row); v2 = __ldg(row + 1); v3 = __ldg(row + 2);
    }
    const float* wr = ws + (b * 4) * AT;
    fma_tile<AT>(acc, v0, wr);
    fma_tile<AT>(acc, v1, wr + AT);
    fma_tile<AT>(acc, v2, wr + 2 * AT);
    fma_tile<AT>(acc, v3, wr + 3 * AT);
  }
#pragma unroll
  for (int i = 0; i < AT; ++i) {
    const int a = a0 + i;
    if (a < A) {
      const int64_t off = (int64_t)a * l + j;
      small_[n * small_ns + off] = apply_epi(acc[i], epi, aux + n * aux_ns, off);
    }
  }
}

// ------------------------------------------------------------------ 1-D up ---------------
template <int BT>
__global__ void __launch_bounds__(CONV_THREADS)
up1d_kernel(const float* __restrict__ small_, int64_t small_ns, const float* __restrict__ w,
            const float* __restrict__ bias, const float* __restrict__ aux, int64_t aux_ns,
            float* __restrict__ big, int64_t big_ns, int64_t N, int A, int Bc, int l, int pad, int epi,
            int vec_ok) {
  extern __shared__ __align__(16) float ws[];  // [A][BT][4]
  const int b0 = blockIdx.y * BT;
  for (int idx = threadIdx.x; idx < A * BT * 4; idx += blockDim.x) {
    const int a = idx / (BT * 4), rem = idx - a * (BT * 4);
    const int b = b0 + (rem >> 2);
    ws[idx] = b < Bc ? w[((int64_t)a * Bc + b0) * 4 + rem] : 0.f;
  }
  __syncthreads();
  const int64_t pix = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= N * l) return;
  const int64_t n = pix / l;
  const int i = (int)(pix - n * l);
  const int64_t Lb = 4 * (int64_t)l;
  float acc[BT * 4];
#pragma unroll
  for (int q = 0; q < BT * 4; ++q) acc[q] = (bias != nullptr && b0 + (q >> 2) < Bc) ? bias[b0 + (q >> 2)] : 0.f;
  const float* sp = small_ + n * small_ns + i;
  for (int a = 0; a < A; ++a) {
    const float v = __ldg(sp + (int64_t)a * l);
    fma_tile<BT * 4>(acc, v, ws + a * (BT * 4));
  }
  const float* auxn = aux + n * aux_ns;
  float* outn = big + n * big_ns;
#pragma unroll
  for (int bi = 0; bi < BT; ++bi) {
    const int b = b0 + bi;
    if (b >= Bc) continue;
    const int64_t base = (int64_t)b * Lb + 4 * (int64_t)i - pad;
    if (pad == 0 && vec_ok && epi != LSHM_EPI_DELU) {
      float4 o;
      o.x = apply_epi(acc[bi * 4 + 0], epi, nullptr, 0);
      o.y = apply_epi(acc[bi * 4 + 1], epi, nullptr, 0);
      o.z = apply_epi(acc[bi * 4 + 2], epi, nullptr, 0);
      o.w = apply_epi(acc[bi * 4 + 3], epi, nullptr, 0);
      *reinterpret_cast<float4*>(outn + base) = o;
    } else {
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const int64_t off = base + t;
        if (pad == 0 || !(i == 0 && t == 0)) outn[off] = apply_epi(acc[bi * 4 + t], epi, auxn, off);
      }
      if (pad == 1 && i == l - 1) {  // last position receives no tap: bias only
        const int64_t off = (int64_t)b * Lb + Lb - 1;
        const float bv = bias != nullptr ? bias[b] : 0.f;
        outn[off] = apply_epi(bv, epi, auxn, off);
      }
    }
  }
}

// ------------------------------------------------------------------ 1-D wgrad ------------
template <int AT>
__global__ void __launch_bounds__(WG_THREADS)
wgrad1d_kernel(const float* __restrict__ small_, int64_t small_ns, const float* __restrict__ big,
               int64_t big_ns, float* __restrict__ dw, int64_t N, int A, int Bc, int l, int pad,
               int64_t chunk) {
  __shared__ float red[WG_THREADS / 32][AT * 4];
  const int combo = blockIdx.y;
  const int b = combo % Bc, a0 = (combo / Bc) * AT;
  const int64_t Lb = 4 * (int64_t)l, total = N * (int64_t)l;
  const int64_t start = (int64_t)blockIdx.x * chunk;
  const int64_t stop = min(start + chunk, total);
  float acc[AT][4];
#pragma unroll
  for (int i = 0; i < AT; ++i)
#pragma unroll
    for (int t = 0; t < 4; ++t) acc[i][t] = 0.f;
  for (int64_t pix = start + threadIdx.x; pix < stop; pix += WG_THREADS) {
    const int64_t n = pix / l;
    const int j = (int)(pix - n * l);
    const float* row = big + n * big_ns + (int64_t)b * Lb + 4 * (int64_t)j - pad;
    float win[4];
    win[0] = (pad == 0 || j > 0) ? __ldg(row) : 0.f;
    win[1] = __ldg(row + 1);
    win[2] = __ldg(row + 2);
    win[3] = __ldg(row + 3);
    const float* sp = small_ + n * small_ns + j;
#pragma unroll
    for (int i = 0; i < AT; ++i) {
      const float s = (a0 + i < A) ? __ldg(sp + (int64_t)(a0 + i) * l) : 0.f;
#pragma unroll
      for (int t = 0; t < 4; ++t) acc[i][t] = fmaf(s, win[t], acc[i][t]);
    }
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < AT; ++i)
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const float v = warp_sum(acc[i][t]);
      if (lane == 0) red[wid][i * 4 + t] = v;
    }
  __syncthreads();
  if (threadIdx.x < AT * 4) {
    float v = 0.f;
#pragma unroll
    for (int q = 0; q < WG_THREADS / 32; ++q) v += red[q][threadIdx.x];
    const int i = threadIdx.x >> 2, t = threadIdx.x & 3;
    if (a0 + i < A) atomicAdd(dw + ((int64_t)(a0 + i) * Bc + b) * 4 + t, v);
  }
}

// ------------------------------------------------------------------ bias gradient --------
__global__ void __launch_bounds__(256)
channel_sum_kernel(const float* __restrict__ g, int64_t g_ns, float* __restrict__ db, int64_t N,
                   int Cn, int64_t len, int64_t chunk) {
  __shared__ float red[32];
  const int c = blockIdx.y;
  const int64_t total = N * len;
  const int64_t start = (int64_t)blockIdx.x * chunk, stop = min(start + chunk, total);
  float s = 0.f;
  for (int64_t idx = start + threadIdx.x; idx < stop; idx += blockDim.x) {
    const int64_t n = idx / len, r = idx - n * len;
    s += __ldg(g + n * g_ns + (int64_t)c * len + r);
  }
  s = block_sum<float>(s, red);
  if (threadIdx.x == 0) atomicAdd(db + c, s);
}

int64_t pick_chunk(int64_t total, int64_t combos, int threads) {
  // aim for ~16 blocks per SM overall, at least 4 and at most 64 pixels per thread
  const int64_t target_blocks = std::max<int64_t>(1, (int64_t)sm_count() * 16 / std::max<int64_t>(1, combos));
  int64_t ppt = ceil_div(total, target_blocks * threads);
  ppt = std::max<int64_t>(4, std::min<int64_t>(64, ppt));
  return ppt * threads;
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace
}  // namespace lshm

using namespace lshm;

#define CHECK_CONV_ARGS(name)                                                                   \
  LSHM_REQUIRE(small_ && big && w, "%s: null pointer", name);                                   \
  LSHM_REQUIRE(N >= 0 && A > 0 && Bc > 0, "%s: bad sizes N=%lld A=%d B=%d", name, (long long)N, A, Bc); \
  LSHM_REQUIRE(epilogue >= 0 && epilogue <= 2, "%s: bad epilogue %d", name, epilogue);          \
  LSHM_REQUIRE(epilogue != LSHM_EPI_DELU || aux != nullptr, "%s: DELU epilogue needs aux", name)

extern "C" {



int lshm_wgrad2d(const float* small_, int64_t small_ns, const float* big, int64_t big_ns,
                 float* dw, int64_t N, int A, int Bc, int h, int w_, lshm_stream_t stream) {
  LSHM_REQUIRE(small_ && big && dw, "lshm_wgrad2d: null pointer");
  LSHM_REQUIRE(N >= 0 && A > 0 && Bc > 0 && h > 0 && w_ > 0, "lshm_wgrad2d: bad sizes");
  cudaStream_t st = as_stream(stream);
  LSHM_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)A * Bc * 16, st), "lshm_wgrad2d");
  if (N == 0) return LSHM_OK;
  constexpr int AT = 4;
  const int64_t combos = ceil_div(A, AT) * Bc, total = N * h * w_;
  const int64_t chunk = pick_chunk(total, combos, WG_THREADS);
  dim3 grid((unsigned)ceil_div(total, chunk), (unsigned)combos);
  wgrad2d_kernel<AT><<<grid, WG_THREADS, 0, st>>>(small_, small_ns, big, big_ns, dw, N, A, Bc, h, w_, chunk);
  LSHM_CHECK_LAUNCH("lshm_wgrad2d");
  return LSHM_OK;
}



int lshm_wgrad1d(const float* small_, int64_t small_ns, const float* big, int64_t big_ns,
                 float* dw, int64_t N, int A, int Bc, int l, int pad, lshm_stream_t stream) {
  LSHM_REQUIRE(small_ && big && dw, "lshm_wgrad1d: null pointer");
  LSHM_REQUIRE(N >= 0 && A > 0 && Bc > 0 && l > 0 && (pad == 0 || pad == 1), "lshm_wgrad1d: bad sizes");
  cudaStream_t st = as_stream(stream);
  LSHM_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)A * Bc * 4, st), "lshm_wgrad1d");
  if (N == 0) return LSHM_OK;
  constexpr int AT = 8;
  const int64_t combos = ceil_div(A, AT) * Bc, total = N * l;
  const int64_t chunk = pick_chunk(total, combos, WG_THREADS);
  dim3 grid((unsigned)ceil_div(total, chunk), (unsigned)combos);
  wgrad1d_kernel<AT><<<grid, WG_THREADS, 0, st>>>(small_, small_ns, big, big_ns, dw, N, A, Bc, l, pad, chunk);
  LSHM_CHECK_LAUNCH("lshm_wgrad1d");
  return LSHM_OK;
}

int lshm_channel_sum(const float* g, int64_t g_ns, float* db, int64_t N, int Cn, int64_t len,
                     lshm_stream_t stream) {
  LSHM_REQUIRE(g && db && N >= 0 && Cn > 0 && len > 0, "lshm_channel_sum: bad arguments");
  cudaStream_t st = as_stream(stream);
  LSHM_CUDA(cudaMemsetAsync(db, 0, sizeof(float) * Cn, st), "lshm_channel_sum");
  if (N == 0) return LSHM_OK;
  const int64_t total = N * len;
  const int64_t chunk = pick_chunk(total, Cn, 256);
  dim3 grid((unsigned)ceil_div(total, chunk), (unsigned)Cn);
  channel_sum_kernel<<<grid, 256, 0, st>>>(g, g_ns, db, N, Cn, len, chunk);
  LSHM_CHECK_LAUNCH("lshm_channel_sum");
  return LSHM_OK;
}

}  // extern "C"
