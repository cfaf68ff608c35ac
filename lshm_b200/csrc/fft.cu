// Fourier-space features: fftshift(fft2_ortho(x - xhat)) -> cat(Re, Im) (or cat(|F|, arg F)) -> clamp.
//
// Reference semantics: /root/reference/Demo.ipynb:169-174 with torch_fftshift,
// /root/reference/src/lofar_tools.py:24-30 (roll by size//2 on dims 2,3).
//
// One CTA owns one 128x128 (patch, channel) plane: the plane is read once from HBM (64 KB, or
// 128 KB with xhat) and the two output planes are written once (128 KB) - the algorithmic
// minimum.  Round-2 structure (profiles/r2_ncu_fft.md: the round-1 kernel was bound by LSU wavefronts,
// 4-8 different 128-byte lines per global load / store instruction):
//  * the plane is staged by the copy engine: 128 bulk copies (cp.async.bulk, one 512-byte row each) into
//    ONE shared-memory buffer of 128 rows at a 136-float pitch, completion on an mbarrier;
//  * everything then happens IN PLACE in that buffer: the row pass reads real rows (paired r, r+64: the input
//    is real, two rows go through one complex FFT), exchanges its two register stages through the same
//    two rows, and leaves the half spectrum (columns 0..63 re | im, Nyquist column in the row's spare floats);
//    the column pass (warp = 8 columns) transforms in place and leaves row u at position rho(u);
//  * the output is then emitted row-contiguously: a half-warp writes 256 consecutive bytes per instruction
//    (st.global.v4), the mirrored half F[-u,-v] = conj F[u,v] is read backwards from the same buffer.
// Each 128-point transform is 16 x 8 (Cooley-Tukey) in registers.  69.6 KB of shared memory per CTA.
#include "tc_common.cuh"

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

// torch.clamp semantics (NaN stays NaN) in two instructions
__device__ __forceinline__ float clampn(float a, float c) {
  float r;
  asm("min.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(c));
  asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(r), "f"(-c));
  return r;
}

constexpr int RS = 136;   // row pitch (floats) of the plane buffer: 128 samples / 64 re + 64 im, Nyquist re, im, 6 spare

// Half spectrum of row r: re at [r][pcol(r,k)], im at [r][64 + pcol(r,k)], k = 0..63.  The buffer is addressed by
// lanes that differ in the column (8 consecutive) and in the row, where the row step is 1 (first column stage),
// 8 (second column stage) or 2 (row-pass stores).  bank = 8*(row + group) + col%8 at a 136-float pitch, so the
// 8-column group is XOR-ed with the row's /8 digit: every one of those access patterns is conflict-free.
__device__ __forceinline__ int pcol(int row, int k) { return k ^ (((row >> 3) & 3) << 3); }
// the column pass leaves row frequency u = k1 + 16*k2 where its second stage computed it: row 8*k1 + brev3(k2)
__device__ __forceinline__ int rho(int u) { return ((u & 15) << 3) | brev3(u >> 4); }

// (the ortho scale 1/128 is already in the values: it rides in the column pass's inter-stage twiddles)
template <int MODE>
__device__ __forceinline__ void emit4(float* __restrict__ o0, float* __restrict__ o1, float4 re, float4 im, float clamp) {
  float4 a, b;
  if (MODE == LSHM_FFT_REIM) {
    a = make_float4(clampn(re.x, clamp), clampn(re.y, clamp), clampn(re.z, clamp), clampn(re.w, clamp));
    b = make_float4(clampn(im.x, clamp), clampn(im.y, clamp), clampn(im.z, clamp), clampn(im.w, clamp));
  } else {
    a = make_float4(clampn(sqrtf(re.x * re.x + im.x * im.x), clamp), clampn(sqrtf(re.y * re.y + im.y * im.y), clamp),
                    clampn(sqrtf(re.z * re.z + im.z * im.z), clamp), clampn(sqrtf(re.w * re.w + im.w * im.w), clamp));
    b = make_float4(atan2f(im.x, re.x), atan2f(im.y, re.y), atan2f(im.z, re.z), atan2f(im.w, re.w));
  }
  *reinterpret_cast<float4*>(o0) = a;
  *reinterpret_cast<float4*>(o1) = b;
}

// Column pass over the half spectrum in the plane buffer, in place.  NYQ = false: lane -> (column 8*warp + lane%8,
// row offset lane/8), four row offsets per iteration; NYQ = true: the Nyquist column (one warp), lane = row offset.
// `tws` is the twiddle table times 1/128 (the ortho scale of the 2-D transform, exact: a power of two).
template <bool NYQ>
__device__ __forceinline__ void column_pass(float* __restrict__ B, const float2* __restrict__ tws, int warp, int lane) {
  constexpr int IM = NYQ ? 1 : 64;                 // offset of the imaginary part
  const int col = NYQ ? 128 : warp * 8 + (lane & 7);
  const int sub = NYQ ? lane : (lane >> 3);
  constexpr int PER_IT = NYQ ? 32 : 4;
  const int cx[4] = {col, NYQ ? col : col ^ 8, NYQ ? col : col ^ 16, NYQ ? col : col ^ 24};   // pcol(r, col) by (r/8)%4
  // stage 1: (col, n2): 16-point FFT over rows 8*n1 + n2
#pragma unroll 1
  for (int it = 0; it * PER_IT < 8; ++it) {
    const int n2 = it * PER_IT + sub;
    if (n2 < 8) {
      float* base = B + n2 * RS;
      cpx v[16];
#pragma unroll
      for (int n1 = 0; n1 < 16; ++n1) {
        const float* p = base + n1 * 8 * RS + cx[n1 & 3];
        v[n1] = cpx{p[0], p[IM]};
      }
      fft_dif<16>(v);
      const float2* twp = tws + n2;                // twiddle exp(-2*pi*i*k1*n2/128)/128 at tws[k1*n2]
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int k1 = brev4(j);
        const float2 w = tws[k1 * n2];
        const cpx y = cmul(v[j], cpx{w.x, w.y});
        float* p = base + k1 * 8 * RS + cx[k1 & 3];
        p[0] = y.r; p[IM] = y.i;
      }
      (void)twp;
    }
  }
  __syncwarp();
  // stage 2: (col, k1): 8-point FFT over n2; register j = frequency k1 + 16*brev3(j) stays in row 8*k1 + j
#pragma unroll 1
  for (int it = 0; it * PER_IT < 16; ++it) {
    const int k1 = it * PER_IT + sub;
    if (k1 < 16) {
      float* base = B + k1 * 8 * RS + (NYQ ? col : (col ^ ((k1 & 3) << 3)));
      cpx v[8];
#pragma unroll
      for (int n2 = 0; n2 < 8; ++n2) v[n2] = cpx{base[n2 * RS], base[n2 * RS + IM]};
      fft_dif<8>(v);
#pragma unroll
      for (int j = 0; j < 8; ++j) { base[j * RS] = v[j].r; base[j * RS + IM] = v[j].i; }
    }
  }
}

template <int MODE>
__global__ void __launch_bounds__(FFT_THREADS, 2)
fft2_kernel(const float* __restrict__ x, const float* __restrict__ xhat, float* __restrict__ out,
            int C, float clamp) {
  extern __shared__ __align__(16) float B[];   // [128][RS]
  __shared__ float2 tw[FN], tws[FN];          // exp(-2*pi*i*t/128), and the same / 128
  __shared__ __align__(8) uint64_t bar;
  const int64_t plane = blockIdx.x;            // n*C + c
  const int64_t n = plane / C;
  const int c = (int)(plane - n * C);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const float* src = x + plane * FN * FN;
  const float* src2 = xhat ? xhat + plane * FN * FN : nullptr;

  // ------------------------------------------------------------------ stage the plane (copy engine), twiddles meanwhile
  if (tid == 0) {
    tc::mbar_init(&bar, 1);
    tc::mbar_init_fence();
    tc::mbar_arrive_expect_tx(&bar, FN * FN * 4);
  }
  __syncthreads();
  if (tid < FN) {
    tc::bulk_g2s(B + tid * RS, src + tid * FN, FN * 4, &bar);
  } else {
    float sn, cs;
    sincospif(-(float)(tid - FN) / 64.f, &sn, &cs);    // exp(-2*pi*i*t/128)
    tw[tid - FN] = make_float2(cs, sn);
    tws[tid - FN] = make_float2(cs * (1.f / 128.f), sn * (1.f / 128.f));
  }
  __syncthreads();
  tc::mbar_wait(&bar, 0);

  // ------------------------------------------------------------------ rows: warp w owns the pairs (r, r + 64), r = 8w..8w+7
  // Real input: rows are transformed in PAIRS (z = a + i*b, one complex 128-point FFT, then
  // A[k] = (Z[k] + conj Z[-k])/2, B[k] = (Z[k] - conj Z[-k])/(2i)), and only columns k = 0..64 are kept.
  cpx vin[16];
  auto load_half = [&](int half) {
    const int pl = lane >> 3, n2 = lane & 7;
    const int ra = warp * 8 + half * 4 + pl;
    const float* pa = B + ra * RS + n2;
    const float* pb = pa + 64 * RS;
#pragma unroll
    for (int n1 = 0; n1 < 16; ++n1) {
      float a = pa[n1 * 8], b = pb[n1 * 8];
      if (src2) { a -= __ldg(src2 + ra * FN + n1 * 8 + n2); b -= __ldg(src2 + (ra + 64) * FN + n1 * 8 + n2); }
      vin[n1] = {a, b};
    }
  };
  load_half(0);
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    {
      // stage 1: lane -> (pair = 8w + 4*half + lane/8, n2 = lane%8): 16-point FFT over n = 8*n1 + n2; the
      // twiddled result replaces the lane's own samples (re in row r, im in row r + 64)
      const int pl = lane >> 3, n2 = lane & 7;
      float* yre = B + (warp * 8 + half * 4 + pl) * RS;
      float* yim = yre + 64 * RS;
      cpx v[16];
#pragma unroll
      for (int n1 = 0; n1 < 16; ++n1) v[n1] = vin[n1];
      fft_dif<16>(v);
#pragma unroll
      for (int j = 0; j < 16; ++j) {               // register j holds k1 = brev4(j)
        const int k1 = brev4(j);
        const int t = k1 * n2;                     // twiddle exp(-2*pi*i*k1*n2/128)
        const float2 w = tw[t];
        const cpx y = cmul(v[j], cpx{w.x, w.y});
        yre[k1 * 8 + n2] = y.r;
        yim[k1 * 8 + n2] = y.i;
      }
    }
    __syncwarp();
    if (half == 0) load_half(1);                   // (xhat comes straight from global: in flight during stage 2)
#pragma unroll
    for (int sub = 0; sub < 2; ++sub) {
      // stage 2: lane -> (pair-in-half = sub + 2*(lane/16), k1 = lane%16): 8-point FFT over n2, then unpack
      const int pl = sub + 2 * (lane >> 4), k1 = lane & 15;
      const int ra = warp * 8 + half * 4 + pl;
      float* rowa = B + ra * RS;
      float* rowb = rowa + 64 * RS;
      cpx v[8];
      {
        // 16-byte loads at a 32-byte lane stride: lanes 4..7 of each quarter-warp fetch their upper
        // half first so the eight lanes of a wavefront cover all 32 banks
        const int sw = (k1 >> 2) & 1;
        const float* yr_ = rowa + k1 * 8;
        const float* yi_ = rowb + k1 * 8;
        const float4 ra0 = *reinterpret_cast<const float4*>(yr_ + 4 * sw);
        const float4 ra1 = *reinterpret_cast<const float4*>(yr_ + 4 * (sw ^ 1));
        const float4 rb0 = *reinterpret_cast<const float4*>(yi_ + 4 * sw);
        const float4 rb1 = *reinterpret_cast<const float4*>(yi_ + 4 * (sw ^ 1));
        const float4 a0 = sw ? ra1 : ra0, a1 = sw ? ra0 : ra1, b0 = sw ? rb1 : rb0, b1 = sw ? rb0 : rb1;
        v[0] = {a0.x, b0.x}; v[1] = {a0.y, b0.y}; v[2] = {a0.z, b0.z}; v[3] = {a0.w, b0.w};
        v[4] = {a1.x, b1.x}; v[5] = {a1.y, b1.y}; v[6] = {a1.z, b1.z}; v[7] = {a1.w, b1.w};
      }
      __syncwarp();                                // the half spectrum overwrites the exchange rows
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
            const int pc = pcol(ra, k);            // rows r and r + 64 share the swizzle
            rowa[pc] = ar; rowa[64 + pc] = ai;
            rowb[pc] = br; rowb[64 + pc] = bi;
          } else {
            rowa[128] = ar; rowa[129] = ai; rowb[128] = br; rowb[129] = bi;
          }
        }
      }
    }
    __syncwarp();
  }
  __syncthreads();

  // ------------------------------------------------------------------ columns 0..64: warp w owns 8w..8w+7, warp 0 also column 64
  column_pass<false>(B, tws, warp, lane);
  if (warp == 0) column_pass<true>(B, tws, warp, lane);
  __syncthreads();

  // ------------------------------------------------------------------ emit: scale, fftshift, clamp; whole rows, 16 bytes per lane
  // Output row ur holds u = ur - 64: columns 64..127 are F[u, 0..63], column 0 is F[u, 64] and columns 1..63 are
  // F[u, -(64 - vc)] = conj F[-u, 64 - vc], read backwards from row -u.  A half-warp writes one half row.
  float* ore = out + ((n * 2 * C + c) * (int64_t)FN) * FN;
  float* oim = out + ((n * 2 * C + C + c) * (int64_t)FN) * FN;
#pragma unroll 2
  for (int i = 0; i < 8; ++i) {
    const int ur = (warp * 8 + i) * 2 + (lane >> 4), l = lane & 15;
    const int u = (ur + 64) & 127;
    const int rr = rho(u), rm = rho((128 - u) & 127);
    {
      const float* p = B + rr * RS + pcol(rr, 4 * l);
      const float4 re = *reinterpret_cast<const float4*>(p), im = *reinterpret_cast<const float4*>(p + 64);
      emit4<MODE>(ore + ur * FN + 64 + 4 * l, oim + ur * FN + 64 + 4 * l, re, im, clamp);
    }
    {
      const float* p = B + rm * RS + pcol(rm, 60 - 4 * l);          // F[-u, 60-4l .. 63-4l]
      const float4 re = *reinterpret_cast<const float4*>(p), im = *reinterpret_cast<const float4*>(p + 64);
      float r0 = __shfl_up_sync(0xffffffffu, re.x, 1), i0 = -__shfl_up_sync(0xffffffffu, im.x, 1);   // F[-u, 64-4l]
      if (l == 0) { r0 = B[rr * RS + 128]; i0 = B[rr * RS + 129]; }                                   // F[u, 64]
      emit4<MODE>(ore + ur * FN + 4 * l, oim + ur * FN + 4 * l, make_float4(r0, re.w, re.z, re.y),
                  make_float4(i0, -im.w, -im.z, -im.y), clamp);
    }
  }
}

template <int MODE>
int launch_fft(const float* x, const float* xhat, float* out, int64_t N, int C, float clamp, cudaStream_t st) {
  const size_t smem = (size_t)FN * RS * sizeof(float);
  LSHM_CUDA(cudaFuncSetAttribute(fft2_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "lshm_fft2");
  fft2_kernel<MODE><<<(unsigned)(N * C), FFT_THREADS, smem, st>>>(x, xhat, out, C, clamp);
  LSHM_CHECK_LAUNCH("lshm_fft2");
  return LSHM_OK;
}

}  // namespace
}  // namespace lshm

using namespace lshm;

extern "C" {

int lshm_fft2_features(const float* x, const float* xhat, float* out,
                       int64_t N, int C, float clamp, int mode, lshm_stream_t stream) {
  LSHM_REQUIRE(x && out && N >= 0 && C > 0, "lshm_fft2_features: bad arguments");
  LSHM_REQUIRE(mode == LSHM_FFT_REIM || mode == LSHM_FFT_MAGPHASE, "lshm_fft2_features: bad mode %d", mode);
  LSHM_REQUIRE(N * C < (1LL << 31), "lshm_fft2_features: too many planes for one launch");
  LSHM_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out)) & 15) == 0,
               "lshm_fft2_features: x and out must be 16-byte aligned");
  if (N == 0) return LSHM_OK;
  if (mode == LSHM_FFT_MAGPHASE) return launch_fft<LSHM_FFT_MAGPHASE>(x, xhat, out, N, C, clamp, as_stream(stream));
  return launch_fft<LSHM_FFT_REIM>(x, xhat, out, N, C, clamp, as_stream(stream));
}

int lshm_fft2_reim_shift_clamp(const float* x, const float* xhat, float* out,
                               int64_t N, int C, float clamp, lshm_stream_t stream) {
  return lshm_fft2_features(x, xhat, out, N, C, clamp, LSHM_FFT_REIM, stream);
}

}  // extern "C"
