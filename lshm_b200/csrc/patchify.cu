// Loader hot path: int8 visibilities x per-(baseline,freq,pol) scale -> half-overlapping
// 128x128 patches, clamp, and the global z-score.
//
// Reference semantics: get_data_minibatch / get_data_for_baseline,
// /root/reference/src/lofar_tools.py:113-141 (scale), :157-173 (unfold + patch-major row order
// n = (ci*py+cj)*nb + k), :187/:333 (clamp), :190-193/:336-338 (mean / unbiased std).
//
// One fused gather: the [.., 4 pol, 2 re/im] int8 octet of a (t,f) sample is exactly the 8
// output channels, so a thread reads four 8-byte octets (32 contiguous bytes) plus four float4
// scale vectors and writes one float4 per channel; a warp covers one 128-wide patch row, i.e.
// 1 KB contiguous in and 512 B contiguous out per channel.  The statistics for the z-score are
// accumulated in the same pass (block reduce in double, one atomic per block), so the
// normalisation needs exactly one more read+write of the patches.
#include "common.cuh"

namespace lshm {
namespace {

__device__ __forceinline__ float clampf(float v, float c) {
  return v != v ? v : fminf(fmaxf(v, -c), c);
}

template <int C>
__global__ void __launch_bounds__(256)
patchify_kernel(const int8_t* __restrict__ vis, const float* __restrict__ scale,
                const int32_t* __restrict__ sel, int nb, int T, int F, int P, int px, int py,
                float clamp, float* __restrict__ y, double* __restrict__ stats) {
  __shared__ double red[32];
  const int P4 = P >> 2, s = P >> 1;
  const int64_t total = (int64_t)nb * px * py * P * P4;
  float lsum = 0.f, lsq = 0.f;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int f4 = (int)(idx % P4);
    int64_t rest = idx / P4;
    const int t = (int)(rest % P);
    const int64_t n = rest / P;
    const int k = (int)(n % nb);
    const int pidx = (int)(n / nb);
    const int ci = pidx / py, cj = pidx - ci * py;
    const int ts = ci * s + t, fs = cj * s + f4 * 4;
    const int b = sel[k];
    float out[C][4];
#pragma unroll
    for (int c = 0; c < C; ++c)
#pragma unroll
      for (int q = 0; q < 4; ++q) out[c][q] = 0.f;
    if (ts < T) {
      const int8_t* vp = vis + (((int64_t)b * T + ts) * F + fs) * 8;
      const float* sp = scale + ((int64_t)b * F + fs) * 4;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (fs + q < F) {
          const uint2 raw = __ldg(reinterpret_cast<const uint2*>(vp + q * 8));
          const float4 sc = __ldg(reinterpret_cast<const float4*>(sp + q * 4));
          const float scv[4] = {sc.x, sc.y, sc.z, sc.w};
          const uint32_t words[2] = {raw.x, raw.y};
#pragma unroll
          for (int c = 0; c < C; ++c) {
            // C=8: byte c (pol=c/2, ri=c%2).  C=4: pols 0 and 3 -> bytes 0,1,6,7.
            const int byte = (C == 8) ? c : (c < 2 ? c : c + 4);
            const int pol = byte >> 1;
            const int8_t v8 = (int8_t)((words[byte >> 2] >> ((byte & 3) * 8)) & 0xff);
            out[c][q] = clampf((float)v8 * scv[pol], clamp);
          }
        }
      }
    }
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const float4 o = make_float4(out[c][0], out[c][1], out[c][2], out[c][3]);
      *reinterpret_cast<float4*>(y + (((n * C + c) * P + t) * (int64_t)P) + f4 * 4) = o;
#pragma unroll
      for (int q = 0; q < 4; ++q) { lsum += out[c][q]; lsq = fmaf(out[c][q], out[c][q], lsq); }
    }
  }
  if (stats != nullptr) {
    const double a = block_sum<double>((double)lsum, red);
    if (threadIdx.x == 0) atomicAdd(stats, a);
    const double q = block_sum<double>((double)lsq, red);
    if (threadIdx.x == 0) atomicAdd(stats + 1, q);
  }
}

__global__ void __launch_bounds__(256)
normalise_kernel(float* __restrict__ y, int64_t n, const double* __restrict__ stats, int64_t n_stats) {
  const double cnt = (double)n_stats;
  const double mean_d = stats[0] / cnt;
  const double var_d = (stats[1] - stats[0] * stats[0] / cnt) / (cnt - 1.0);
  const float mean = (float)mean_d, sd = (float)sqrt(var_d);
  const int64_t n4 = n >> 2;
  float4* y4 = reinterpret_cast<float4*>(y);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4;
       i += (int64_t)gridDim.x * blockDim.x) {
    float4 v = y4[i];
    v.x = (v.x - mean) / sd; v.y = (v.y - mean) / sd; v.z = (v.z - mean) / sd; v.w = (v.w - mean) / sd;
    y4[i] = v;
  }
  if (blockIdx.x == 0)
    for (int64_t i = (n4 << 2) + threadIdx.x; i < n; i += blockDim.x) y[i] = (y[i] - mean) / sd;
}

}  // namespace
}  // namespace lshm

using namespace lshm;

extern "C" {

int lshm_patchify_scale_i8(const int8_t* vis, const float* scale, const int32_t* sel,
                           int nb, int T, int F, int C, int P, float clamp,
                           float* y, double* stats, lshm_stream_t stream) {
  LSHM_REQUIRE(vis && scale && sel && y, "lshm_patchify_scale_i8: null pointer");
  LSHM_REQUIRE(C == 4 || C == 8, "lshm_patchify_scale_i8: num_channels must be 4 or 8 (got %d)", C);
  LSHM_REQUIRE(nb >= 0 && T > 0 && F > 0 && P >= 8 && (P & 7) == 0, "lshm_patchify_scale_i8: bad sizes");
  LSHM_REQUIRE((reinterpret_cast<uintptr_t>(vis) & 7) == 0 && (reinterpret_cast<uintptr_t>(scale) & 15) == 0 &&
               (reinterpret_cast<uintptr_t>(y) & 15) == 0, "lshm_patchify_scale_i8: misaligned buffer");
  if (nb == 0) return LSHM_OK;
  const int Tp = T > P ? T : P, Fp = F > P ? F : P, s = P / 2;
  const int px = (Tp - P) / s + 1, py = (Fp - P) / s + 1;
  const int64_t total = (int64_t)nb * px * py * P * (P / 4);
  const int64_t blocks = std::min<int64_t>(ceil_div(total, 256), (int64_t)sm_count() * 16);
  if (C == 8)
    patchify_kernel<8><<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(vis, scale, sel, nb, T, F, P, px, py, clamp, y, stats);
  else
    patchify_kernel<4><<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(vis, scale, sel, nb, T, F, P, px, py, clamp, y, stats);
  LSHM_CHECK_LAUNCH("lshm_patchify_scale_i8");
  return LSHM_OK;
}

int lshm_normalise(float* y, int64_t n, const double* stats, lshm_stream_t stream) {
  return lshm_normalise_n(y, n, stats, n, stream);
}

int lshm_normalise_n(float* y, int64_t n, const double* stats, int64_t n_stats, lshm_stream_t stream) {
  LSHM_REQUIRE(y && stats && n >= 0 && n_stats >= 2 && n_stats >= n, "lshm_normalise: bad arguments");
  if (n == 0) return LSHM_OK;
  LSHM_REQUIRE((reinterpret_cast<uintptr_t>(y) & 15) == 0, "lshm_normalise: y must be 16-byte aligned");
  const int64_t blocks = std::min<int64_t>(ceil_div(n >> 2, 256) + 1, (int64_t)sm_count() * 16);
  normalise_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(y, n, stats, n_stats);
  LSHM_CHECK_LAUNCH("lshm_normalise");
  return LSHM_OK;
}

}  // extern "C"
