// Fourier-space features: fftshift(fft2_ortho(x - xhat)) -> cat(Re, Im) -> clamp.
//
// Reference semantics: /root/reference/Demo.ipynb:169-174 with torch_fftshift,
// /root/reference/src/lofar_tools.py:24-30 (roll by size//2 on dims 2,3).
//
// One CTA owns one 128x128 (patch, channel) plane: the plane is read once from HBM (64 KB, or
// 128 KB with xhat) and the two output planes are written once (128 KB) - the algorithmic
// minimum.  Each 128-point transform is 16 x 8 (Cooley-Tukey): a 16-point FFT in registers over
// the stride-8 samples, the inter-stage twiddle, an exchange through shared memory, an 8-point
// FFT in registers.  A warp owns 16 rows for the whole row pass and 16 columns for the whole
// column pass, so the two register stages of a pass are separated by __syncwarp only; the block
// synchronises once between the passes.  The first stage reads its samples straight from global
// memory (8 lanes = one 32-byte sector), the last stage applies the ortho scale, the fftshift and
// the clamp and stores rows coalesced (lanes = consecutive columns).
#include "common.cuh"

namespace lshm {
namespace {

constexpr int FN = 128, FFT_THREADS = 256, FS = 144;   // FS: padded row stride (floats): +16 banks per row

struct cpx { float r, i; };
__device__ __forceinline__ cpx cmul(cpx a, cpx b) { return {a.r * b.r - a.i * b.i, a.r * b.i + a.i * b.r}; }

// in-register radix-2 DIF FFT of length N (8 or 16); output in bit-reversed order
template <int N>
__device__ __forceinline__ void fft_dif(cpx (&v)[N]) {
  constexpr float C8 = 0.70710678118654752f;
  constexpr float C16a = 0.92387953251128674f, C16b = 0.38268343236508977f;
#pragma unroll
  for (int h = N / 2; h >= 1; h >>= 1) {
#pragma unroll
    for (int i = 0; i < N; ++i) {
      if ((i & h) == 0) {
        const int k = i & (h - 1);                       // twiddle exp(-2*pi*i*k/(2h))
        const cpx a = v[i], b = v[i + h];
        v[i] = {a.r + b.r, a.i + b.i};
        const cpx d = {a.r - b.r, a.i - b.i};
        const int num = k * (N / (2 * h));              // exp(-2*pi*i*num/N)
        cpx w;
        // constant twiddles for N <= 16
        const int idx16 = num * (16 / N);
        switch (idx16) {
          case 0: w = {1.f, 0.f}; break;
          case 1: w = {C16a, -C16b}; break;
          case 2: w = {C8, -C8}; break;
          case 3: w = {C16b, -C16a}; break;
          case 4: w = {0.f, -1.f}; break;
          case 5: w = {-C16b, -C16a}; break;
          case 6: w = {-C8, -C8}; break;
          default: w = {-C16a, -C16b}; break;
        }
        if (idx16 == 0) v[i + h] = d;
        else if (idx16 == 4) v[i + h] = {d.i, -d.r};
        else v[i + h] = cmul(d, w);
      }
    }
  }
}

__device__ __forceinline__ constexpr int brev3(int v) { return ((v & 1) << 2) | (v & 2) | ((v >> 2) & 1); }
__device__ __forceinline__ constexpr int brev4(int v) { return ((v & 1) << 3) | ((v & 2) << 1) | ((v >> 1) & 2) | ((v >> 3) & 1); }

__device__ __forceinline__ float clampn(float a, float c) { return a != a ? a : fminf(fmaxf(a, -c), c); }

__global__ void __launch_bounds__(FFT_THREADS)
fft2_kernel(const float* __restrict__ x, const float* __restrict__ xhat, float* __restrict__ out,
            int C, float clamp) {
  extern __shared__ __align__(16) float sm[];
  float* re = sm;                 // [128][FS]
  float* im = sm + FN * FS;       // [128][FS]
  __shared__ float twr[FN], twi[FN];
  const int64_t plane = blockIdx.x;           // n*C + c
  const int64_t n = plane / C;
  const int c = (int)(plane - n * C);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid < FN) {
    float sn, cs;
    sincospif(-(float)tid / 64.f, &sn, &cs);  // exp(-2*pi*i*tid/128)
    twr[tid] = cs; twi[tid] = sn;
  }
  __syncthreads();
  const float* src = x + plane * FN * FN;
  const float* src2 = xhat ? xhat + plane * FN * FN : nullptr;

  // ------------------------------------------------------------------ rows: warp w owns rows 16w..16w+15
  for (int it = 0; it < 4; ++it) {
    // stage 1: lane -> (row = 16w + 4it + lane/8, n2 = lane%8): 16-point FFT over n = 8*n1 + n2
    const int r = warp * 16 + it * 4 + (lane >> 3), n2 = lane & 7;
    cpx v[16];
#pragma unroll
    for (int n1 = 0; n1 < 16; ++n1) {
      float a = __ldg(src + r * FN + n1 * 8 + n2);
      if (src2) a -= __ldg(src2 + r * FN + n1 * 8 + n2);
      v[n1] = {a, 0.f};
    }
    fft_dif<16>(v);
#pragma unroll
    for (int j = 0; j < 16; ++j) {                 // register j holds k1 = brev4(j)
      const int k1 = brev4(j);
      const int t = k1 * n2;                       // twiddle exp(-2*pi*i*k1*n2/128)
      const cpx y = cmul(v[j], cpx{twr[t], twi[t]});
      re[r * FS + k1 * 8 + n2] = y.r;
      im[r * FS + k1 * 8 + n2] = y.i;
    }
  }
  __syncwarp();
  for (int it = 0; it < 8; ++it) {
    // stage 2: lane -> (row = 16w + 2it + lane/16, k1 = lane%16): 8-point FFT over n2
    const int r = warp * 16 + it * 2 + (lane >> 4), k1 = lane & 15;
    cpx v[8];
    {
      const float4 a0 = *reinterpret_cast<const float4*>(re + r * FS + k1 * 8);
      const float4 a1 = *reinterpret_cast<const float4*>(re + r * FS + k1 * 8 + 4);
      const float4 b0 = *reinterpret_cast<const float4*>(im + r * FS + k1 * 8);
      const float4 b1 = *reinterpret_cast<const float4*>(im + r * FS + k1 * 8 + 4);
      v[0] = {a0.x, b0.x}; v[1] = {a0.y, b0.y}; v[2] = {a0.z, b0.z}; v[3] = {a0.w, b0.w};
      v[4] = {a1.x, b1.x}; v[5] = {a1.y, b1.y}; v[6] = {a1.z, b1.z}; v[7] = {a1.w, b1.w};
    }
    fft_dif<8>(v);
    __syncwarp();                                  // every lane has read its inputs of these two rows
#pragma unroll
    for (int j = 0; j < 8; ++j) {                  // register j holds k2 = brev3(j); k = k1 + 16*k2
      const int k = k1 + 16 * brev3(j);
      re[r * FS + k] = v[j].r;
      im[r * FS + k] = v[j].i;
    }
  }
  __syncthreads();

  // ------------------------------------------------------------------ columns: warp w owns columns 16w..16w+15
  for (int it = 0; it < 4; ++it) {
    // stage 1: lane -> (col = 16w + lane%16, n2 = 2it + lane/16): 16-point FFT over rows 8*n1 + n2, in place
    const int col = warp * 16 + (lane & 15), n2 = it * 2 + (lane >> 4);
    cpx v[16];
#pragma unroll
    for (int n1 = 0; n1 < 16; ++n1) v[n1] = {re[(n1 * 8 + n2) * FS + col], im[(n1 * 8 + n2) * FS + col]};
    fft_dif<16>(v);
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const int k1 = brev4(j);
      const int t = k1 * n2;
      const cpx y = cmul(v[j], cpx{twr[t], twi[t]});
      re[(k1 * 8 + n2) * FS + col] = y.r;
      im[(k1 * 8 + n2) * FS + col] = y.i;
    }
  }
  __syncwarp();
  float* ore = out + ((n * 2 * C + c) * (int64_t)FN) * FN;
  float* oim = out + ((n * 2 * C + C + c) * (int64_t)FN) * FN;
  const float sc = 1.f / 128.f;
  for (int it = 0; it < 8; ++it) {
    // stage 2: lane -> (col = 16w + lane%16, k1 = 2it + lane/16): 8-point FFT over n2, then store
    const int col = warp * 16 + (lane & 15), k1 = it * 2 + (lane >> 4);
    cpx v[8];
#pragma unroll
    for (int n2 = 0; n2 < 8; ++n2) v[n2] = {re[(k1 * 8 + n2) * FS + col], im[(k1 * 8 + n2) * FS + col]};
    fft_dif<8>(v);
    const int vcol = (col + 64) & 127;             // fftshift
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int u = k1 + 16 * brev3(j);            // row frequency index
      const int urow = (u + 64) & 127;
      ore[urow * FN + vcol] = clampn(v[j].r * sc, clamp);
      oim[urow * FN + vcol] = clampn(v[j].i * sc, clamp);
    }
  }
}

}  // namespace
}  // namespace lshm

using namespace lshm;

extern "C" {

int lshm_fft2_reim_shift_clamp(const float* x, const float* xhat, float* out,
                               int64_t N, int C, float clamp, lshm_stream_t stream) {
  LSHM_REQUIRE(x && out && N >= 0 && C > 0, "lshm_fft2_reim_shift_clamp: bad arguments");
  LSHM_REQUIRE(N * C < (1LL << 31), "lshm_fft2_reim_shift_clamp: too many planes for one launch");
  if (N == 0) return LSHM_OK;
  const size_t smem = 2 * FN * FS * sizeof(float);
  LSHM_CUDA(cudaFuncSetAttribute(fft2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
            "lshm_fft2_reim_shift_clamp");
  fft2_kernel<<<(unsigned)(N * C), FFT_THREADS, smem, as_stream(stream)>>>(x, xhat, out, C, clamp);
  LSHM_CHECK_LAUNCH("lshm_fft2_reim_shift_clamp");
  return LSHM_OK;
}

}  // extern "C"
