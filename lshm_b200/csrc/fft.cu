// Fourier-space features: fftshift(fft2_ortho(x - xhat)) -> cat(Re, Im) -> clamp.
//
// Reference semantics: /root/reference/Demo.ipynb:169-174 with torch_fftshift,
// /root/reference/src/lofar_tools.py:24-30 (roll by size//2 on dims 2,3).
//
// One CTA owns one 128x128 (patch, channel) plane: the plane is read once from HBM (64 KB, or
// 128 KB with xhat) and the two output planes are written once (128 KB) - the algorithmic
// minimum.  Each 128-point transform is 16 x 8 (Cooley-Tukey): a 16-point FFT in registers over
// the stride-8 samples, the inter-stage twiddle, an exchange through shared memory, an 8-point
// FFT in registers.  A warp owns 16 rows for the whole row pass and 8 columns for the whole
// column pass, so the two register stages of a pass are separated by __syncwarp only; the block
// synchronises once between the passes.  The input is real: rows go through the complex FFT in
// pairs and only the half spectrum (columns 0..64) is computed, the rest is its conjugate mirror.  The first stage reads its samples straight from global
// memory (8 lanes = one 32-byte sector), the last stage applies the ortho scale, the fftshift and
// the clamp and stores rows coalesced (lanes = consecutive columns).
#include "common.cuh"

namespace lshm {
namespace {

constexpr int FN = 128, FFT_THREADS = 256;

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

constexpr int YS = 136;   // row stride (floats) of the per-warp stage-1 -> stage-2 exchange buffer
constexpr int XS = 64;    // row stride (floats) of the half-spectrum X[128 rows][columns 0..63]; column 64 apart

// X is addressed by lanes that differ in the column (8 consecutive) and in the row, where the row
// step is 1 (first column stage: n2), 8 (second column stage: k1) or 2 (row-pass stores).  With any
// padded row stride one of the three collides (a 72-float stride made every second-stage column load
// a 4-way bank conflict: 37% extra shared-memory wavefronts in the ncu capture), so the 8-column
// group is XOR-swizzled with two row bits taken from both the row's low and its /8 digits.
__device__ __forceinline__ int xidx(int row, int col) { return row * XS + (col ^ (((row ^ (row >> 3)) & 3) << 3)); }

// Real input: rows are transformed in PAIRS (z = a + i*b, one complex 128-point FFT, then
// A[k] = (Z[k] + conj Z[-k])/2, B[k] = (Z[k] - conj Z[-k])/(2i)), and only columns k = 0..64 are kept:
// the other half of the 2-D spectrum is the conjugate mirror F[-u,-v] = conj F[u,v], written from
// the same registers.  Half the butterflies, half the shared memory (two CTAs per SM).
__global__ void __launch_bounds__(FFT_THREADS, 2)
fft2_kernel(const float* __restrict__ x, const float* __restrict__ xhat, float* __restrict__ out,
            int C, float clamp) {
  extern __shared__ __align__(16) float sm[];
  float* xre = sm;                          // [128][XS] swizzled (xidx)
  float* xim = xre + FN * XS;               // [128][XS]
  float* ybase = xim + FN * XS;             // per warp: re[4][YS], im[4][YS]
  __shared__ float twr[FN], twi[FN];
  __shared__ float nre[FN], nim[FN];        // the Nyquist column (k = 64) of the row transforms
  const int64_t plane = blockIdx.x;         // n*C + c
  const int64_t n = plane / C;
  const int c = (int)(plane - n * C);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  float* yre = ybase + warp * (8 * YS);
  float* yim = yre + 4 * YS;
  const float* src = x + plane * FN * FN;
  const float* src2 = xhat ? xhat + plane * FN * FN : nullptr;

  // ------------------------------------------------------------------ rows: warp w owns row pairs 8w..8w+7
  // The samples of the second half are requested right after the first half's register stage, so their
  // latency is covered by the first half's second stage (the load phase was 23 % of the stall samples).
  cpx vin[16];
  auto load_half = [&](int half) {
    const int pl = lane >> 3, n2 = lane & 7;
    const int ra = 2 * (warp * 8 + half * 4 + pl);
#pragma unroll
    for (int n1 = 0; n1 < 16; ++n1) {
      float a = __ldg(src + ra * FN + n1 * 8 + n2), b = __ldg(src + (ra + 1) * FN + n1 * 8 + n2);
      if (src2) { a -= __ldg(src2 + ra * FN + n1 * 8 + n2); b -= __ldg(src2 + (ra + 1) * FN + n1 * 8 + n2); }
      vin[n1] = {a, b};
    }
  };
  load_half(0);                                 // in flight while the twiddle table is built
  if (tid < FN) {
    float sn, cs;
    sincospif(-(float)tid / 64.f, &sn, &cs);    // exp(-2*pi*i*tid/128)
    twr[tid] = cs; twi[tid] = sn;
  }
  __syncthreads();
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    {
      // stage 1: lane -> (pair = 8w + 4*half + lane/8, n2 = lane%8): 16-point FFT over n = 8*n1 + n2
      const int pl = lane >> 3, n2 = lane & 7;
      cpx v[16];
#pragma unroll
      for (int n1 = 0; n1 < 16; ++n1) v[n1] = vin[n1];
      fft_dif<16>(v);
#pragma unroll
      for (int j = 0; j < 16; ++j) {               // register j holds k1 = brev4(j)
        const int k1 = brev4(j);
        const int t = k1 * n2;                     // twiddle exp(-2*pi*i*k1*n2/128)
        const cpx y = cmul(v[j], cpx{twr[t], twi[t]});
        yre[pl * YS + k1 * 8 + n2] = y.r;
        yim[pl * YS + k1 * 8 + n2] = y.i;
      }
    }
    __syncwarp();
    if (half == 0) load_half(1);
    for (int sub = 0; sub < 2; ++sub) {
      // stage 2: lane -> (pair-in-half = 2*sub + lane/16, k1 = lane%16): 8-point FFT over n2, then unpack
      const int pl = 2 * sub + (lane >> 4), k1 = lane & 15;
      const int ra = 2 * (warp * 8 + half * 4 + pl);
      cpx v[8];
      {
        // 16-byte loads at a 32-byte lane stride: lanes 4..7 of each quarter-warp fetch their upper
        // half first so the eight lanes of a wavefront cover all 32 banks
        const int sw = (k1 >> 2) & 1;
        const float* yr_ = yre + pl * YS + k1 * 8;
        const float* yi_ = yim + pl * YS + k1 * 8;
        const float4 ra0 = *reinterpret_cast<const float4*>(yr_ + 4 * sw);
        const float4 ra1 = *reinterpret_cast<const float4*>(yr_ + 4 * (sw ^ 1));
        const float4 rb0 = *reinterpret_cast<const float4*>(yi_ + 4 * sw);
        const float4 rb1 = *reinterpret_cast<const float4*>(yi_ + 4 * (sw ^ 1));
        const float4 a0 = sw ? ra1 : ra0, a1 = sw ? ra0 : ra1, b0 = sw ? rb1 : rb0, b1 = sw ? rb0 : rb1;
        v[0] = {a0.x, b0.x}; v[1] = {a0.y, b0.y}; v[2] = {a0.z, b0.z}; v[3] = {a0.w, b0.w};
        v[4] = {a1.x, b1.x}; v[5] = {a1.y, b1.y}; v[6] = {a1.z, b1.z}; v[7] = {a1.w, b1.w};
      }
      fft_dif<8>(v);                               // register j holds Z[k1 + 16*brev3(j)]
      const int srcl = (lane & 16) | ((16 - k1) & 15);   // lane holding k1' = 16 - k1 of the same pair
#pragma unroll
      for (int k2 = 0; k2 < 5; ++k2) {
        // partner Z[128-k]: (k1' = 16-k1, k2' = 7-k2) for k1 != 0, (0, (8-k2)%8) for k1 == 0
        const int jo = brev3(k2), ja = brev3((7 - k2) & 7), jb = brev3((8 - k2) & 7);
        float pr = 0.f, pi = 0.f;
        if (k2 < 4) {
          pr = __shfl_sync(0xffffffffu, v[ja].r, srcl);
          pi = __shfl_sync(0xffffffffu, v[ja].i, srcl);
        }
        if (k1 == 0) { pr = v[jb].r; pi = v[jb].i; }
        if (k2 < 4 || k1 == 0) {
          const int k = k1 + 16 * k2;
          const float zr = v[jo].r, zi = v[jo].i;
          const float ar = 0.5f * (zr + pr), ai = 0.5f * (zi - pi);     // A[k] = (Z[k] + conj Zp)/2
          const float br = 0.5f * (zi + pi), bi = -0.5f * (zr - pr);    // B[k] = -i/2 (Z[k] - conj Zp)
          if (k2 < 4) {
            xre[xidx(ra, k)] = ar; xim[xidx(ra, k)] = ai;
            xre[xidx(ra + 1, k)] = br; xim[xidx(ra + 1, k)] = bi;
          } else {
            nre[ra] = ar; nim[ra] = ai; nre[ra + 1] = br; nim[ra + 1] = bi;
          }
        }
      }
    }
    __syncwarp();
  }
  __syncthreads();

  // ------------------------------------------------------------------ columns 0..64: warp w owns 8w..8w+7, warp 0 also column 64
  float* ore = out + ((n * 2 * C + c) * (int64_t)FN) * FN;
  float* oim = out + ((n * 2 * C + C + c) * (int64_t)FN) * FN;
  const float sc = 1.f / 128.f;
  for (int grp = 0; grp < 2; ++grp) {
    if (grp == 1 && warp != 0) break;              // the Nyquist column
    const int ncol = grp == 0 ? 8 : 1;
    const int cl = grp == 0 ? (lane & 7) : 0;
    const int sub = grp == 0 ? (lane >> 3) : lane; // row-offset index within an iteration
    const int per_it = grp == 0 ? 4 : 32;
    const int col = grp == 0 ? warp * 8 + cl : 64;
    (void)ncol;
    // stage 1: (col, n2): 16-point FFT over rows 8*n1 + n2, in place
    for (int it = 0; it * per_it < 8; ++it) {
      const int n2 = it * per_it + sub;
      if (n2 < 8) {
        cpx v[16];
#pragma unroll
        for (int n1 = 0; n1 < 16; ++n1) {
          const int r = n1 * 8 + n2;
          v[n1] = grp == 0 ? cpx{xre[xidx(r, col)], xim[xidx(r, col)]} : cpx{nre[r], nim[r]};
        }
        fft_dif<16>(v);
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int k1 = brev4(j);
          const int t = k1 * n2;
          const cpx y = cmul(v[j], cpx{twr[t], twi[t]});
          const int r = k1 * 8 + n2;
          if (grp == 0) { xre[xidx(r, col)] = y.r; xim[xidx(r, col)] = y.i; }
          else { nre[r] = y.r; nim[r] = y.i; }
        }
      }
    }
    __syncwarp();
    // stage 2: (col, k1): 8-point FFT over n2, then the two mirrored stores
    for (int it = 0; it * per_it < 16; ++it) {
      const int k1 = it * per_it + sub;
      if (k1 < 16) {
        cpx v[8];
#pragma unroll
        for (int n2 = 0; n2 < 8; ++n2) {
          const int r = k1 * 8 + n2;
          v[n2] = grp == 0 ? cpx{xre[xidx(r, col)], xim[xidx(r, col)]} : cpx{nre[r], nim[r]};
        }
        fft_dif<8>(v);
        const int vc = (col + 64) & 127;             // fftshift of column v = col
        const int vm = (128 - col + 64) & 127;       // ... and of the mirrored column -v
        const bool mirror = col >= 1 && col <= 63;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int u = k1 + 16 * brev3(j);          // row frequency index
          const float fr = clampn(v[j].r * sc, clamp), fi = clampn(v[j].i * sc, clamp);
          const int ur = (u + 64) & 127;
          ore[ur * FN + vc] = fr;
          oim[ur * FN + vc] = fi;
          if (mirror) {
            const int um = ((128 - u) + 64) & 127;   // row of -u after the shift
            ore[um * FN + vm] = fr;
            oim[um * FN + vm] = clampn(-v[j].i * sc, clamp);
          }
        }
      }
    }
    __syncwarp();
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
  const size_t smem = (2 * FN * XS + 8 * 8 * YS) * sizeof(float);
  LSHM_CUDA(cudaFuncSetAttribute(fft2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
            "lshm_fft2_reim_shift_clamp");
  fft2_kernel<<<(unsigned)(N * C), FFT_THREADS, smem, as_stream(stream)>>>(x, xhat, out, C, clamp);
  LSHM_CHECK_LAUNCH("lshm_fft2_reim_shift_clamp");
  return LSHM_OK;
}

}  // extern "C"
