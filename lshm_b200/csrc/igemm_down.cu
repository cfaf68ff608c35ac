// "down" implicit GEMM on tcgen05: Conv2d(k4,s2,p1) / Conv1d(k4,s4,pad) forward and the dgrad of the
// matching transposed convs.  Reference semantics: F.conv2d / F.conv1d at
// /root/reference/src/lofar_models.py:73-78,:158-163 (and the autograd of :93-98,:178-183).
//
// Formulation (DESIGN.md "conv kernels"): space-to-depth turns the k4/s2 conv into a k2/s1 conv over
// Z[q, c] with c = 4*b + (sy*2+sx) and q a position of the zero-padded (h+1)x(w+1) block grid:
//     out[q, a] = sum_{tap=(ty,tx)} sum_c Z[q + ty*(w+1) + tx, c] * Wt[tap][a][c]
// (1-D, k4/s4: one tap, Z[q, 4*b+t] = big[b, 4j-pad+t]).  The CTA stages ONE bf16 hi/lo copy of its
// Z tile in shared memory (rows at a uniform 16-byte pitch) and every tap is the same tile read
// through a descriptor whose start address is shifted by the tap offset - no im2col duplication.
// Weights arrive pre-split / pre-permuted ("image", lshm_conv_prep) by one bulk copy per K block.
// Warp roles: 4 producer+epilogue warps, 1 MMA-issuing warp, 1 weight-loading warp; a ring of
// shared-memory stages (mbarrier full/empty), fp32 accumulators in TMEM, bias/ELU/ELU' fused in
// the epilogue which writes NCHW / NCL directly.
#include "tc_common.cuh"

namespace lshm {
namespace {

using namespace tc;

struct DownArgs {
  const float* big; int64_t big_ns;
  const uint8_t* wimg;
  const float* bias;
  const float* aux; int64_t aux_ns;
  float* small_; int64_t small_ns;
  int64_t N; int A; int Bc; int h; int w; int pad; int epi;
  int slots; int nstage; int64_t Q;
};

template <int DIM, int NT, int KC>
__global__ void __launch_bounds__(192) igemm_down_kernel(DownArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t full_bar[4], empty_bar[4], acc_bar;
  __shared__ uint32_t tmem_base;
  constexpr int T = DIM == 2 ? 4 : 1;
  constexpr int CC = KC / 8;
  constexpr uint32_t IMG = 2u * T * CC * NT * 16;
  constexpr uint32_t TMEM_COLS = NT <= 32 ? 32 : (NT <= 64 ? 64 : 128);
  const int SLOTS = a.slots, NS = a.nstage;
  const uint32_t zbytes = (uint32_t)CC * SLOTS * 16;
  const uint32_t stage_bytes = 2 * zbytes + IMG;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t q0 = (int64_t)blockIdx.x * 128;
  const int nt = blockIdx.y;
  const int Kc = 4 * a.Bc;
  const int KB = (Kc + KC - 1) / KC;
  const int PW = a.w + 1, PH = a.h + 1;

  if (warp == 4) tmem_alloc(&tmem_base, TMEM_COLS);
  if (tid == 0) {
    for (int s = 0; s < 4; ++s) { mbar_init(&full_bar[s], 5); mbar_init(&empty_bar[s], 1); }
    mbar_init(&acc_bar, 1);
    mbar_init_fence();
  }
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem = tmem_base;

  if (warp < 4) {
    // ------------------------------------------------ producers: stage the Z tile (hi/lo bf16)
    // decode this thread's slots once
    int64_t sn[2]; int sy_[2], sx_[2]; bool sv[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int s = tid + i * 128;
      const int64_t q = q0 + s;
      sv[i] = s < SLOTS && q < a.Q;
      sn[i] = 0; sy_[i] = 0; sx_[i] = 0;
      if (sv[i]) {
        if (DIM == 2) {
          const int64_t pp = (int64_t)PH * PW;
          sn[i] = q / pp;
          const int r = (int)(q - sn[i] * pp);
          sy_[i] = r / PW; sx_[i] = r - sy_[i] * PW;
        } else {
          sn[i] = q / a.w;                      // 1-D: a.w holds the small length l
          sx_[i] = (int)(q - sn[i] * a.w);
        }
      }
    }
    const int H = 2 * a.h, W = 2 * a.w;
    for (int kb = 0; kb < KB; ++kb) {
      const int s = kb % NS, ph = (kb / NS) & 1;
      mbar_wait(&empty_bar[s], ph ^ 1);
      uint8_t* zhi = smem + (size_t)s * stage_bytes;
      uint8_t* zlo = zhi + zbytes;
      const int ccb = (min(KC, Kc - kb * KC)) >> 3;
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int slot = tid + i * 128;
        if (slot >= SLOTS) continue;
        // all chunk columns of this slot are fetched before any is converted (loads in flight)
        float v[CC][8];
#pragma unroll
        for (int cc = 0; cc < CC; ++cc) {
#pragma unroll
          for (int e = 0; e < 8; ++e) v[cc][e] = 0.f;
          if (cc < ccb && sv[i]) {
            const int b0 = 2 * (kb * CC + cc);
#pragma unroll
            for (int bb = 0; bb < 2; ++bb) {
              const int b = b0 + bb;
              if (b < a.Bc) {
                if (DIM == 2) {
                  const float* base = a.big + sn[i] * a.big_ns + (int64_t)b * H * W;
                  const int r0 = 2 * sy_[i] - 1, c0 = 2 * sx_[i] - 1;
#pragma unroll
                  for (int yy = 0; yy < 2; ++yy) {
                    const int r = r0 + yy;
                    const bool rin = r >= 0 && r < H;
#pragma unroll
                    for (int xx = 0; xx < 2; ++xx) {
                      const int c = c0 + xx;
                      if (rin && c >= 0 && c < W) v[cc][bb * 4 + yy * 2 + xx] = __ldg(base + (int64_t)r * W + c);
                    }
                  }
                } else {
                  const int64_t Lb = 4 * (int64_t)a.w;
                  const float* base = a.big + sn[i] * a.big_ns + (int64_t)b * Lb + 4 * (int64_t)sx_[i] - a.pad;
                  if (a.pad == 0) {
                    const float4 q4 = __ldg(reinterpret_cast<const float4*>(base));
                    v[cc][bb * 4 + 0] = q4.x; v[cc][bb * 4 + 1] = q4.y; v[cc][bb * 4 + 2] = q4.z; v[cc][bb * 4 + 3] = q4.w;
                  } else {
#pragma unroll
                    for (int t = 0; t < 4; ++t)
                      if (t > 0 || sx_[i] > 0) v[cc][bb * 4 + t] = __ldg(base + t);
                  }
                }
              }
            }
          }
        }
#pragma unroll
        for (int cc = 0; cc < CC; ++cc) {
          if (cc < ccb) {
            uint4 hi, lo;
            split8(v[cc], hi, lo);
            *reinterpret_cast<uint4*>(zhi + ((size_t)cc * SLOTS + slot) * 16) = hi;
            *reinterpret_cast<uint4*>(zlo + ((size_t)cc * SLOTS + slot) * 16) = lo;
          }
        }
      }
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&full_bar[s]);
    }
    // ------------------------------------------------ epilogue: TMEM -> bias/act -> global
    mbar_wait(&acc_bar, 0);
    fence_after();
    const bool ok = sv[0] && (DIM == 1 || (sy_[0] < a.h && sx_[0] < a.w));
    const int64_t hw = DIM == 2 ? (int64_t)a.h * a.w : (int64_t)a.w;
    const int64_t pos = DIM == 2 ? (int64_t)sy_[0] * a.w + sx_[0] : (int64_t)sx_[0];
    float* outp = a.small_ + sn[0] * a.small_ns + pos;
    const float* auxp = a.aux != nullptr ? a.aux + sn[0] * a.aux_ns + pos : nullptr;
#pragma unroll 1
    for (int g = 0; g < NT / 16; ++g) {
      float v[16];
      tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + g * 16, v);
      if (ok) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int ch = nt * NT + g * 16 + j;
          if (ch < a.A) {
            float r = v[j] + (a.bias != nullptr ? __ldg(a.bias + ch) : 0.f);
            if (a.epi == LSHM_EPI_ELU) r = elu_f(r);
            else if (a.epi == LSHM_EPI_DELU) r *= delu_from_out(__ldg(auxp + ch * hw));
            outp[ch * hw] = r;
          }
        }
      }
    }
  } else if (warp == 4) {
    // ------------------------------------------------ MMA issuer (one elected lane)
    if (lane == 0) {
      const uint32_t idesc = make_idesc(NT, 0, 0);
      uint32_t acc = 0;
      for (int kb = 0; kb < KB; ++kb) {
        const int s = kb % NS, ph = (kb / NS) & 1;
        mbar_wait(&full_bar[s], ph);
        fence_after();
        const uint32_t zhi = smem_u32(smem + (size_t)s * stage_bytes);
        const uint32_t zlo = zhi + zbytes;
        const uint32_t bhi = zlo + zbytes;
        const uint32_t blo = bhi + IMG / 2;
        const int ksteps = (min(KC, Kc - kb * KC)) >> 4;
#pragma unroll
        for (int tap = 0; tap < T; ++tap) {
          const uint32_t shift = DIM == 2 ? (uint32_t)((tap >> 1) * PW + (tap & 1)) : 0u;
          for (int ks = 0; ks < ksteps; ++ks) {
            const uint32_t aoff = ((uint32_t)(2 * ks) * SLOTS + shift) * 16;
            const uint32_t boff = ((uint32_t)(tap * CC + 2 * ks) * NT) * 16;
            mma_split3(tmem, make_desc(zhi + aoff, SLOTS * 16, 128), make_desc(zlo + aoff, SLOTS * 16, 128),
                       make_desc(bhi + boff, NT * 16, 128), make_desc(blo + boff, NT * 16, 128), idesc, acc);
            acc = 1;
          }
        }
        commit(&empty_bar[s]);
      }
      commit(&acc_bar);
    }
  } else {
    // ------------------------------------------------ weight image loader
    if (lane == 0) {
      for (int kb = 0; kb < KB; ++kb) {
        const int s = kb % NS, ph = (kb / NS) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        mbar_arrive_expect_tx(&full_bar[s], IMG);
        bulk_g2s(smem + (size_t)s * stage_bytes + 2 * zbytes, a.wimg + ((size_t)nt * KB + kb) * IMG, IMG, &full_bar[s]);
      }
    }
  }
  fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem, TMEM_COLS);
}

// ---------------------------------------------------------------------------------------------
// weight image:  [ntile][kblock][half hi/lo][tap][chunk][a_local][8 x bf16]
// ---------------------------------------------------------------------------------------------
struct DownGeom { int NT, KC, ntiles, KB, T; size_t img; };

DownGeom down_geom(int dim, int A, int Bc) {
  DownGeom g;
  const int a16 = (A + 15) / 16 * 16;
  g.NT = a16 <= 16 ? 16 : (a16 <= 32 ? 32 : (a16 <= 48 ? 48 : 96));
  g.KC = (dim == 2 && g.NT == 96) ? 16 : 32;
  g.ntiles = (A + g.NT - 1) / g.NT;
  g.KB = (4 * Bc + g.KC - 1) / g.KC;
  g.T = dim == 2 ? 4 : 1;
  g.img = (size_t)2 * g.T * (g.KC / 8) * g.NT * 16;
  return g;
}

__global__ void prep_down_kernel(const float* __restrict__ w, int dim, int A, int Bc, int NT, int KC, int KB, int T,
                                 int64_t total, uint8_t* __restrict__ img) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // one 8-element chunk
  if (idx >= total) return;
  const int CC = KC / 8;
  int64_t r = idx;
  const int al = (int)(r % NT); r /= NT;
  const int cc = (int)(r % CC); r /= CC;
  const int tap = (int)(r % T); r /= T;
  const int kb = (int)(r % KB); r /= KB;
  const int nt = (int)r;
  const int a = nt * NT + al;
  float v[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int c = kb * KC + cc * 8 + e;
    const int b = c >> 2, sub = c & 3;
    float x = 0.f;
    if (a < A && b < Bc) {
      if (dim == 2) {
        const int ky = 2 * (tap >> 1) + (sub >> 1), kx = 2 * (tap & 1) + (sub & 1);
        x = w[(((int64_t)a * Bc + b) * 4 + ky) * 4 + kx];
      } else {
        x = w[((int64_t)a * Bc + b) * 4 + sub];
      }
    }
    v[e] = x;
  }
  uint4 hi, lo;
  tc::split8(v, hi, lo);
  const size_t blk = (size_t)2 * T * CC * NT * 16;
  uint8_t* base = img + ((size_t)nt * KB + kb) * blk + (((size_t)tap * CC + cc) * NT + al) * 16;
  *reinterpret_cast<uint4*>(base) = hi;
  *reinterpret_cast<uint4*>(base + blk / 2) = lo;
}

template <int DIM, int NT, int KC>
int launch_down_t(const DownArgs& a, int ntiles, cudaStream_t st) {
  constexpr int T = DIM == 2 ? 4 : 1;
  const uint32_t zbytes = (uint32_t)(KC / 8) * a.slots * 16;
  const size_t stage = 2 * (size_t)zbytes + (size_t)2 * T * (KC / 8) * NT * 16;
  const size_t smem = stage * a.nstage;
  LSHM_CUDA(cudaFuncSetAttribute(igemm_down_kernel<DIM, NT, KC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "igemm_down");
  dim3 grid((unsigned)ceil_div(a.Q, 128), (unsigned)ntiles);
  igemm_down_kernel<DIM, NT, KC><<<grid, 192, smem, st>>>(a);
  LSHM_CHECK_LAUNCH("igemm_down");
  return LSHM_OK;
}

int launch_down(int dim, DownArgs a, cudaStream_t st) {
  const DownGeom g = down_geom(dim, a.A, a.Bc);
  a.slots = dim == 2 ? (128 + a.w + 2 + 7) / 8 * 8 : 128;
  a.Q = dim == 2 ? a.N * (int64_t)(a.h + 1) * (a.w + 1) : a.N * (int64_t)a.w;
  const size_t stage = (size_t)2 * (g.KC / 8) * a.slots * 16 + g.img;
  int ns = (int)std::min<size_t>(4, std::max<size_t>(1, (96 * 1024) / stage));
  if (ns < 2 && 2 * stage <= 200 * 1024) ns = 2;
  a.nstage = std::min(ns, g.KB);
#define LD(D, NTV, KCV) return launch_down_t<D, NTV, KCV>(a, g.ntiles, st)
  if (dim == 2) {
    switch (g.NT) { case 16: LD(2, 16, 32); case 32: LD(2, 32, 32); case 48: LD(2, 48, 32); default: LD(2, 96, 16); }
  } else {
    switch (g.NT) { case 16: LD(1, 16, 32); case 32: LD(1, 32, 32); case 48: LD(1, 48, 32); default: LD(1, 96, 32); }
  }
#undef LD
}

}  // namespace

// shared with igemm_up.cu / igemm_wgrad.cu through conv_tc.h-style forward declarations
size_t down_image_bytes(int dim, int A, int Bc) {
  const DownGeom g = down_geom(dim, A, Bc);
  return g.img * g.ntiles * g.KB;
}

int prep_down_image(const float* w, int dim, int A, int Bc, void* img, cudaStream_t st) {
  const DownGeom g = down_geom(dim, A, Bc);
  const int64_t total = (int64_t)g.ntiles * g.KB * g.T * (g.KC / 8) * g.NT;
  prep_down_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, st>>>(w, dim, A, Bc, g.NT, g.KC, g.KB, g.T, total,
                                                                    reinterpret_cast<uint8_t*>(img));
  LSHM_CHECK_LAUNCH("lshm_conv_prep(down)");
  return LSHM_OK;
}

}  // namespace lshm

using namespace lshm;

extern "C" {

int lshm_down2d(const float* big, int64_t big_ns, const void* wimg, const float* bias,
                const float* aux, int64_t aux_ns, float* small_, int64_t small_ns,
                int64_t N, int A, int Bc, int h, int w_, int epilogue, lshm_stream_t stream) {
  LSHM_REQUIRE(big && wimg && small_, "lshm_down2d: null pointer");
  LSHM_REQUIRE(N >= 0 && A > 0 && Bc > 0 && (Bc & 3) == 0 && h > 0 && w_ > 0, "lshm_down2d: bad sizes (Bc must be a multiple of 4)");
  LSHM_REQUIRE(epilogue >= 0 && epilogue <= 2, "lshm_down2d: bad epilogue %d", epilogue);
  LSHM_REQUIRE(epilogue != LSHM_EPI_DELU || aux != nullptr, "lshm_down2d: DELU epilogue needs aux");
  LSHM_REQUIRE(w_ <= 512, "lshm_down2d: small-map width %d too large", w_);
  LSHM_REQUIRE((reinterpret_cast<uintptr_t>(wimg) & 15) == 0, "lshm_down2d: weight image must be 16-byte aligned");
  if (N == 0) return LSHM_OK;
  DownArgs a{};
  a.big = big; a.big_ns = big_ns; a.wimg = reinterpret_cast<const uint8_t*>(wimg); a.bias = bias;
  a.aux = epilogue == LSHM_EPI_DELU ? aux : nullptr; a.aux_ns = aux_ns;
  a.small_ = small_; a.small_ns = small_ns; a.N = N; a.A = A; a.Bc = Bc; a.h = h; a.w = w_; a.pad = 0; a.epi = epilogue;
  return launch_down(2, a, as_stream(stream));
}

int lshm_down1d(const float* big, int64_t big_ns, const void* wimg, const float* bias,
                const float* aux, int64_t aux_ns, float* small_, int64_t small_ns,
                int64_t N, int A, int Bc, int l, int pad, int epilogue, lshm_stream_t stream) {
  LSHM_REQUIRE(big && wimg && small_, "lshm_down1d: null pointer");
  LSHM_REQUIRE(N >= 0 && A > 0 && Bc > 0 && (Bc & 3) == 0 && l > 0 && (pad == 0 || pad == 1), "lshm_down1d: bad sizes (Bc must be a multiple of 4)");
  LSHM_REQUIRE(epilogue >= 0 && epilogue <= 2, "lshm_down1d: bad epilogue %d", epilogue);
  LSHM_REQUIRE(epilogue != LSHM_EPI_DELU || aux != nullptr, "lshm_down1d: DELU epilogue needs aux");
  LSHM_REQUIRE((reinterpret_cast<uintptr_t>(wimg) & 15) == 0, "lshm_down1d: weight image must be 16-byte aligned");
  LSHM_REQUIRE(pad == 1 || ((reinterpret_cast<uintptr_t>(big) & 15) == 0 && (big_ns & 3) == 0),
               "lshm_down1d: input must be 16-byte aligned for pad=0");
  if (N == 0) return LSHM_OK;
  DownArgs a{};
  a.big = big; a.big_ns = big_ns; a.wimg = reinterpret_cast<const uint8_t*>(wimg); a.bias = bias;
  a.aux = epilogue == LSHM_EPI_DELU ? aux : nullptr; a.aux_ns = aux_ns;
  a.small_ = small_; a.small_ns = small_ns; a.N = N; a.A = A; a.Bc = Bc; a.h = 1; a.w = l; a.pad = pad; a.epi = epilogue;
  return launch_down(1, a, as_stream(stream));
}

}  // extern "C"
