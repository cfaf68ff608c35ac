// Fourier-space features: fftshift(fft2_ortho(x - xhat)) -> cat(Re, Im) -> clamp.
//
// Reference semantics: /root/reference/Demo.ipynb:169-174 with torch_fftshift,
// /root/reference/src/lofar_tools.py:24-30 (roll by size//2 on dims 2,3).
//
// One CTA owns one 128x128 (patch, channel) plane, entirely in shared memory (2 x 64 KB
// re/im): the plane is read once from HBM (64 KB, or 128 KB with xhat) and the two output planes
// are written once (128 KB) -- the algorithmic minimum.  Decimation-in-frequency radix-2 passes
// run in place along rows then columns (lanes always sweep the contiguous index, so shared
// memory stays conflict-free); the bit-reversed result order and the fftshift are folded into
// the address of the final coalesced store.
#include "common.cuh"

namespace lshm {
namespace {

constexpr int FN = 128, FLOG = 7, FFT_THREADS = 256;

__device__ __forceinline__ int brev7(int v) { return (int)(__brev((unsigned)v) >> 25); }

__global__ void __launch_bounds__(FFT_THREADS)
fft2_kernel(const float* __restrict__ x, const float* __restrict__ xhat, float* __restrict__ out,
            int C, float clamp) {
  extern __shared__ __align__(16) float sm[];
  float* re = sm;               // [128][128]
  float* im = sm + FN * FN;     // [128][128]
  __shared__ float twr[FN / 2], twi[FN / 2];
  const int64_t plane = blockIdx.x;           // n*C + c
  const int64_t n = plane / C;
  const int c = (int)(plane - n * C);
  const int tid = threadIdx.x;
  if (tid < FN / 2) {
    float sn, cs;
    sincospif(-(float)tid / 64.f, &sn, &cs);  // exp(-2*pi*i*tid/128)
    twr[tid] = cs; twi[tid] = sn;
  }
  // ---- load (x - xhat), float4 coalesced
  {
    const float4* src = reinterpret_cast<const float4*>(x + plane * FN * FN);
    const float4* src2 = xhat ? reinterpret_cast<const float4*>(xhat + plane * FN * FN) : nullptr;
    float4* dre = reinterpret_cast<float4*>(re);
    float4* dim = reinterpret_cast<float4*>(im);
    for (int i = tid; i < FN * FN / 4; i += FFT_THREADS) {
      float4 v = ld_nc_f4(reinterpret_cast<const float*>(src + i));
      if (src2) {
        const float4 h = ld_nc_f4(reinterpret_cast<const float*>(src2 + i));
        v.x -= h.x; v.y -= h.y; v.z -= h.z; v.w -= h.w;
      }
      dre[i] = v;
      dim[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  __syncthreads();
  // ---- rows: butterflies (i, i+h) inside each row; lanes sweep the column index
  for (int h = FN / 2; h >= 1; h >>= 1) {
    const int tstep = (FN / 2) / h;
    for (int q = tid; q < FN * (FN / 2); q += FFT_THREADS) {
      const int r = q >> 6, j = q & 63;
      const int k = j & (h - 1);
      const int i0 = r * FN + ((j - k) << 1) + k, i1 = i0 + h;
      const float ar = re[i0], ai = im[i0], br = re[i1], bi = im[i1];
      const float dr = ar - br, di = ai - bi;
      const float wr = twr[k * tstep], wi = twi[k * tstep];
      re[i0] = ar + br; im[i0] = ai + bi;
      re[i1] = dr * wr - di * wi; im[i1] = dr * wi + di * wr;
    }
    __syncthreads();
  }
  // ---- columns: butterflies (row i, row i+h); lanes sweep the column index
  for (int h = FN / 2; h >= 1; h >>= 1) {
    const int tstep = (FN / 2) / h;
    for (int q = tid; q < FN * (FN / 2); q += FFT_THREADS) {
      const int j = q >> 7, col = q & 127;
      const int k = j & (h - 1);
      const int r0 = ((j - k) << 1) + k;
      const int i0 = r0 * FN + col, i1 = i0 + h * FN;
      const float ar = re[i0], ai = im[i0], br = re[i1], bi = im[i1];
      const float dr = ar - br, di = ai - bi;
      const float wr = twr[k * tstep], wi = twi[k * tstep];
      re[i0] = ar + br; im[i0] = ai + bi;
      re[i1] = dr * wr - di * wi; im[i1] = dr * wi + di * wr;
    }
    __syncthreads();
  }
  // ---- store: out[u,v] = F[(u+64)%128, (v+64)%128] / 128, F[k] sits at bit-reversed position
  float* ore = out + ((n * 2 * C + c) * (int64_t)FN) * FN;
  float* oim = out + ((n * 2 * C + C + c) * (int64_t)FN) * FN;
  const float sc = 1.f / 128.f;
  for (int i = tid; i < FN * FN; i += FFT_THREADS) {
    const int u = i >> 7, v = i & 127;
    const int src = brev7((u + 64) & 127) * FN + brev7((v + 64) & 127);
    const float a = re[src] * sc, b = im[src] * sc;   // NaN propagates like Tensor.clamp_
    ore[i] = a != a ? a : fminf(fmaxf(a, -clamp), clamp);
    oim[i] = b != b ? b : fminf(fmaxf(b, -clamp), clamp);
  }
}

}  // namespace
}  // namespace lshm

using namespace lshm;

extern "C" {

int lshm_fft2_reim_shift_clamp(const float* x, const float* xhat, float* out,
                               int64_t N, int C, float clamp, lshm_stream_t stream) {
  LSHM_REQUIRE(x && out && N >= 0 && C > 0, "lshm_fft2_reim_shift_clamp: bad arguments");
  LSHM_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(xhat) & 15) == 0,
               "lshm_fft2_reim_shift_clamp: inputs must be 16-byte aligned");
  LSHM_REQUIRE(N * C < (1LL << 31), "lshm_fft2_reim_shift_clamp: too many planes for one launch");
  if (N == 0) return LSHM_OK;
  const size_t smem = 2 * FN * FN * sizeof(float);
  LSHM_CUDA(cudaFuncSetAttribute(fft2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
            "lshm_fft2_reim_shift_clamp");
  fft2_kernel<<<(unsigned)(N * C), FFT_THREADS, smem, as_stream(stream)>>>(x, xhat, out, C, clamp);
  LSHM_CHECK_LAUNCH("lshm_fft2_reim_shift_clamp");
  return LSHM_OK;
}

}  // extern "C"
