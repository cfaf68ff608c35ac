// Weight-gradient implicit GEMM on tcgen05 for the k4/s2 2-D and k4/s4 1-D (transposed) convolutions:
//     dW[a, b, ky, kx] = sum_{n,oy,ox} S[n,a,oy,ox] * B[n,b,2oy-1+ky,2ox-1+kx]
// (autograd of F.conv2d/conv_transpose2d/conv1d/conv_transpose1d,
//  /root/reference/src/lofar_models.py:73-78,:93-98,:158-163,:178-183).
//
// With the space-to-depth view Z[q, c=4b+sub] of the big map (see igemm_down.cu) this is, per tap,
//     D_tap[a, c] = sum_q S[q, a] * Z[q + shift(tap), c]          (contraction over POSITIONS q)
// so both operands are "MN-major": the staged tiles hold positions as rows at a 16-byte pitch with
// 8 channels per 16 bytes, exactly what the producers of the down/up kernels write; the four taps
// are again four descriptor start addresses into one Z tile and accumulate in four TMEM column
// ranges.  The position range is split over CTAs (split-K); partial results are combined with
// fp32 atomics into dW (zeroed first), already in the reference weight layout.
#include <stdlib.h>
#include "conv_geom.cuh"
#include "tma.cuh"

namespace lshm {
namespace {

using namespace tc;


struct WgArgs {
  // operand planes of the big map (tma.cuh): tensor maps over the hi / lo halves; PRE instances only
  alignas(64) CUtensorMap tm_hi;
  alignas(64) CUtensorMap tm_lo;
  const float* small_; int64_t small_ns;
  const float* big; int64_t big_ns;
  float* dw;
  // FUSED (1-D, operand planes): the data gradient of the same transposed conv from the same staged tiles
  const uint8_t* dimg; float* dz; int64_t dz_ns; uint32_t dimg_off;
  int64_t N; int A; int Bc; int h; int w; int pad;
  int zslots; int nstage; int64_t Q; int64_t kblocks; int64_t kb_per_cta; int ntiles; int scols; int vec_ok;
  int acols;                        // FOLD: chunk columns (8 channels each) per tap, scols = 4 * acols
  FastDiv d_pp, d_pw, d_w, d_zs;   // divisors (h+1)(w+1), w+1, w, zslots
};

// KP = positions per K block: 128 for the shallow layers (fewer pipeline hand-offs per byte), 64 for the
// deep ones (their two tiles are wide and would not fit twice in shared memory).
// Eight producer warps: the kernel is bound by the latency of the gathering loads (38% of the stall
// samples with four), so the bytes in flight per CTA are what matters; warps 0-3 also run the epilogue.
constexpr int WG_NPW = 8;
constexpr int WG_PT = WG_NPW * 32;                 // producer threads
constexpr int WG_THREADS = WG_PT + 32;             // + the MMA warp
// warp roles: producers [0, NPW), (FUSED) data-gradient epilogue [NPW, NPW + 4), then the MMA warp.  The fused 1-D
// instance stages one small tile per K block: four producer warps (the other four become the epilogue); the fused 2-D
// instance stages the small map four times (taps folded into M) and the fp32-input instances gather the big-map tile as
// well: all eight, plus four epilogue warps.
__host__ __device__ constexpr int wg_npw(int dim, bool fused, bool pre) { return (fused && dim == 1 && pre) ? 4 : WG_NPW; }
__host__ __device__ constexpr int wg_mma_warp(int dim, bool fused, bool pre) { return wg_npw(dim, fused, pre) + (fused ? 4 : 0); }
__host__ __device__ constexpr int wg_threads(int dim, bool fused, bool pre) { return (wg_mma_warp(dim, fused, pre) + 1) * 32; }
// PRE: the big map arrives as operand planes; its tile is one tensor-TMA box per half (issued by thread 0),
// the producer warps stage the (8x smaller) small-map tile only.
// FOLD (2-D, A <= 32): the four taps are folded into the M dimension.  D_tap[a,c] = sum_q S[q,a] Z[q+shift_tap,c]
// = sum_q' S[q'-shift_tap, a] Z[q', c]: the small-map tile is staged FOUR times, shifted by the tap offsets, as four
// groups of chunk columns, and one accumulator [(tap,a), c] replaces the four per-tap ones - 24 instead of 96 MMAs
// per 128 positions (the un-folded first layer ran the tensor pipe at 98 %: every MMA costs max(M,128)*N/256 cycles
// whatever the number of real rows, profiles/r2_ncu_planes.md), and the big-map tile needs no halo.
// FUSED (1-D, PRE): the kernel also computes the DATA gradient of the same layer - dz[q, a] = ELU'(S[q, a]) *
// sum_c Z[q, c] Wt[a, c] - from the Z tile it has staged for the weight gradient (the same bytes read as a K-major
// operand, igemm_down.cu) and the small-map values it loads anyway: the gradient planes (the largest tensor of the
// layer) are read from HBM once instead of twice.  Warps 0-3 stage S, warps 4-7 drain the data-gradient accumulator of
// every K block (two TMEM buffers), the MMA warp issues both products.
template <int DIM, int NT, int KP, bool PRE, bool FOLD, bool FUSED>
__global__ void __launch_bounds__(wg_threads(DIM, FUSED, PRE), (KP == 128 && !(FUSED && (DIM == 2 || !PRE)) ? 3 : 2)) igemm_wgrad_kernel(const __grid_constant__ WgArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t full_bar[4], empty_bar[4], acc_bar, dacc_full[2], dacc_empty[2], dimg_bar;
  __shared__ uint32_t tmem_base;
  static_assert(!FUSED || (KP == 128 && (DIM == 1 ? !FOLD : (FOLD == PRE))), "fused data gradient: KP = 128; 2-D: folded plane instance or un-folded fp32 instance");
  constexpr int T = (DIM == 2 && !FOLD) ? 4 : 1;      // accumulators / descriptor shifts per K block
  constexpr int CZ = NT / 8;
  constexpr int NTD = 16;                              // FUSED: channel tile of the data gradient (A <= 16)
  constexpr uint32_t DCOL = T * NT;                    // FUSED: first TMEM column of its two accumulators
  constexpr int TD = DIM == 2 ? 4 : 1;                 // FUSED: taps of the data gradient
  constexpr uint32_t DIMG = 2u * TD * 4 * NTD * 16;    // FUSED: its weight image (lshm_conv_prep "down": hi | lo, [tap][4 chunk columns][16])
  constexpr int NPW = wg_npw(DIM, FUSED, PRE);         // producer warps
  constexpr int MMAW = wg_mma_warp(DIM, FUSED, PRE);   // the MMA-issuing warp
  constexpr int NPT = NPW * 32;
  constexpr uint32_t TCOLS = T * NT + (FUSED ? 2 * NTD : 0);
  constexpr uint32_t TMEM_COLS = TCOLS <= 32 ? 32 : (TCOLS <= 64 ? 64 : (TCOLS <= 128 ? 128 : (TCOLS <= 256 ? 256 : 512)));
  const int ZS = a.zslots, NS = a.nstage;
  // S tile: only the chunk columns that hold data are staged.  The M=128 MMA still walks 16 column
  // groups; the extra groups read bytes that follow the tile (inside this CTA's allocation, see the
  // slack added by the launcher) and produce rows >= A of the accumulator, which nobody reads.
  const uint32_t SBYTES = (uint32_t)a.scols * KP * 16;
  const uint32_t zbytes = (uint32_t)CZ * ZS * 16;
  const uint32_t stage_bytes = 2 * SBYTES + 2 * zbytes;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int mt = blockIdx.y / a.ntiles, nt = blockIdx.y % a.ntiles;
  const int64_t kb0 = (int64_t)blockIdx.x * a.kb_per_cta;
  const int64_t kb1 = min(kb0 + a.kb_per_cta, a.kblocks);
  const int nkb = (int)(kb1 - kb0);
  const int PW = a.w + 1, PH = a.h + 1;
  const int a0 = mt * 128, c0 = nt * NT;
  const int sch = FOLD ? a.scols : min(a.scols, (a.A - a0 + 7) / 8);     // S chunk columns that hold data in this M tile

  if (warp == MMAW) tmem_alloc(&tmem_base, TMEM_COLS);
  if (tid == 0) {
    for (int s = 0; s < 4; ++s) { mbar_init(&full_bar[s], NPW + (PRE ? 1 : 0)); mbar_init(&empty_bar[s], 1); }
    mbar_init(&acc_bar, 1);
    for (int b = 0; b < 2; ++b) { mbar_init(&dacc_full[b], 1); mbar_init(&dacc_empty[b], 4); }
    mbar_init(&dimg_bar, 1);
    mbar_init_fence();
    if (FUSED) {      // the data gradient's weight image stays resident behind the stage ring
      mbar_arrive_expect_tx(&dimg_bar, DIMG);
      bulk_g2s(smem + a.dimg_off, a.dimg, DIMG, &dimg_bar);
    }
  }
  fence_async_smem();
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem = tmem_base;

  if (warp < NPW) {
    const int W = 2 * a.w;
    const int64_t HW2 = 4 * (int64_t)a.h * a.w;     // big-map plane size
    const int64_t hw = DIM == 2 ? (int64_t)a.h * a.w : (int64_t)a.w;
    Ring ring{0, 0};
    for (int it = 0; it < nkb; ++it, ring.next(NS)) {
      const int s = ring.s;
      const int64_t p0 = (kb0 + it) * KP;
      mbar_wait(&empty_bar[s], ring.ph ^ 1);
      uint8_t* shi = smem + (size_t)s * stage_bytes;
      uint8_t* slo = shi + SBYTES;
      uint8_t* zhi = slo + SBYTES;
      uint8_t* zlo = zhi + zbytes;
      if (PRE && tid == 0) {
        mbar_arrive_expect_tx(&full_bar[s], 2 * zbytes);
        tma_load_3d(zhi, &a.tm_hi, 0, (int)(p0 / PLANE_ROW), c0 >> 3, &full_bar[s]);
        tma_load_3d(zlo, &a.tm_lo, 0, (int)(p0 / PLANE_ROW), c0 >> 3, &full_bar[s]);
      }
      const uint32_t up0 = (uint32_t)p0;              // Q < 2^31 (checked by the launcher): 32-bit index math
      const uint32_t uQ = (uint32_t)a.Q, uPW = (uint32_t)PW, upp = (uint32_t)(PH * PW), uw = (uint32_t)a.w;
      // ---- S tile: item = (position, chunk of 8 channels); U items are fetched before any is converted
      constexpr int U = 2;
      for (int item0 = tid; item0 < KP * a.scols; item0 += NPT * U) {
        float v[U][8];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int item = item0 + u * NPT;
          const int p = item % KP, ca = item / KP;
          // FOLD: chunk column ca = tap * acols + channel chunk; this copy is shifted by the tap offset
          const int tap = FOLD ? ca / a.acols : 0;
          const int cch = FOLD ? ca - tap * a.acols : ca;
          const uint32_t shift = FOLD ? (uint32_t)((tap >> 1) * PW + (tap & 1)) : 0u;
          const uint32_t q = up0 + p - shift;             // wraps for the first positions of the map: q >= uQ below
#pragma unroll
          for (int e = 0; e < 8; ++e) v[u][e] = 0.f;
          if (item < KP * a.scols && ca < sch && up0 + p >= shift && q < uQ) {
            const float* sp = nullptr;
            if (DIM == 2) {
              const uint32_t n = fdiv(q, a.d_pp), r = q - n * upp;
              const uint32_t m = fdiv(r, a.d_pw), x = r - m * uPW;
              if (m < (uint32_t)a.h && x < uw) sp = a.small_ + (int64_t)n * a.small_ns + (int64_t)m * a.w + x;
            } else {
              const uint32_t n = fdiv(q, a.d_w);
              sp = a.small_ + (int64_t)n * a.small_ns + (q - n * uw);
            }
            if (sp != nullptr) {
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                const int ch = a0 + cch * 8 + e;
                if (ch < a.A) v[u][e] = __ldg(sp + (int64_t)ch * hw);
              }
            }
          }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int item = item0 + u * NPT;
          if (item < KP * a.scols) {
            const int p = item % KP, ca = item / KP;
            uint4 hi, lo;
            split8(v[u], hi, lo);
            *reinterpret_cast<uint4*>(shi + ((size_t)ca * KP + p) * 16) = hi;
            *reinterpret_cast<uint4*>(slo + ((size_t)ca * KP + p) * 16) = lo;
          }
        }
      }
      // ---- Z tile: item = (slot, chunk of 8 s2d channels)
      constexpr int UZ = 4;
      for (int item0 = tid; !PRE && item0 < ZS * CZ; item0 += NPT * UZ) {
        float v[UZ][8];
#pragma unroll
        for (int u = 0; u < UZ; ++u) {
          const int item = item0 + u * NPT;
          const int cz = (int)fdiv((uint32_t)item, a.d_zs), slot = item - cz * ZS;
          const uint32_t q = up0 + slot;
          const int b0 = (c0 + cz * 8) >> 2;
#pragma unroll
          for (int e = 0; e < 8; ++e) v[u][e] = 0.f;
          if (item < ZS * CZ && q < uQ) {
            if (DIM == 2) {
              const uint32_t n = fdiv(q, a.d_pp), r = q - n * upp;
              const int by = (int)fdiv(r, a.d_pw), bx = (int)r - by * PW;
              // one base pointer and four edge flags per slot; interior slots take the unpredicated path
              const bool r0ok = by > 0, r1ok = by < a.h, c0ok = bx > 0, c1ok = bx < a.w;
              const float* pb = a.big + (int64_t)n * a.big_ns + (int64_t)b0 * HW2 + (int64_t)(2 * by - 1) * W + (2 * bx - 1);
#pragma unroll
              for (int bb = 0; bb < 2; ++bb) {
                if (b0 + bb < a.Bc) {
                  if (r0ok && r1ok && c0ok && c1ok) {
                    v[u][bb * 4 + 0] = __ldg(pb); v[u][bb * 4 + 1] = __ldg(pb + 1);
                    v[u][bb * 4 + 2] = __ldg(pb + W); v[u][bb * 4 + 3] = __ldg(pb + W + 1);
                  } else {
                    if (r0ok && c0ok) v[u][bb * 4 + 0] = __ldg(pb);
                    if (r0ok && c1ok) v[u][bb * 4 + 1] = __ldg(pb + 1);
                    if (r1ok && c0ok) v[u][bb * 4 + 2] = __ldg(pb + W);
                    if (r1ok && c1ok) v[u][bb * 4 + 3] = __ldg(pb + W + 1);
                  }
                }
                pb += HW2;
              }
            } else {
              const uint32_t n = fdiv(q, a.d_w);
              const int j = (int)(q - n * uw);
              const int64_t Lb = 4 * (int64_t)a.w;
#pragma unroll
              for (int bb = 0; bb < 2; ++bb) {
                const int b = b0 + bb;
                if (b < a.Bc) {
                  const float* base = a.big + (int64_t)n * a.big_ns + (int64_t)b * Lb + 4 * (int64_t)j - a.pad;
                  if (a.pad == 0 && a.vec_ok) {
                    const float4 q4 = __ldg(reinterpret_cast<const float4*>(base));
                    v[u][bb * 4 + 0] = q4.x; v[u][bb * 4 + 1] = q4.y; v[u][bb * 4 + 2] = q4.z; v[u][bb * 4 + 3] = q4.w;
                  } else {
#pragma unroll
                    for (int t = 0; t < 4; ++t)
                      if (a.pad == 0 || t > 0 || j > 0) v[u][bb * 4 + t] = __ldg(base + t);
                  }
                }
              }
            }
          }
        }
#pragma unroll
        for (int u = 0; u < UZ; ++u) {
          const int item = item0 + u * NPT;
          if (item < ZS * CZ) {
            const int cz = (int)fdiv((uint32_t)item, a.d_zs), slot = item - cz * ZS;
            uint4 hi, lo;
            split8(v[u], hi, lo);
            *reinterpret_cast<uint4*>(zhi + ((size_t)cz * ZS + slot) * 16) = hi;
            *reinterpret_cast<uint4*>(zlo + ((size_t)cz * ZS + slot) * 16) = lo;
          }
        }
      }
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&full_bar[s]);
    }
    // ------------------------------------------------ epilogue: scatter-add into dW (TMEM lanes = warps 0-3)
    if (warp < 4) {
    mbar_wait(&acc_bar, 0);
    fence_after();
    const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
    if (FOLD) {
      // accumulator row = tap * 8 acols + channel.  M = 64: row 16 j + i sits in TMEM lane 32 j + i (i < 16); M = 128: row = lane
      const int rows = 32 * a.acols;
      const bool m64 = rows <= 64;
      const int row = m64 ? (warp * 16 + lane) : tid;
      const bool rok = (m64 ? lane < 16 : true) && row < rows && nkb > 0;
      const int tapr = row / (8 * a.acols), chn = row - tapr * 8 * a.acols;
      const int ty = tapr >> 1, tx = tapr & 1;
      const bool v2ok = (reinterpret_cast<uintptr_t>(a.dw) & 7) == 0;
#pragma unroll 1
      for (int g = 0; g < NT / 16; ++g) {
        float v[16];
        tmem_ld16(trow + g * 16, v);
        if (rok && chn < a.A) {
#pragma unroll
          for (int jb = 0; jb < 4; ++jb) {
            const int b = (c0 + g * 16) / 4 + jb;
            if (b < a.Bc) {
              // columns 4 b + (sy, sx): kernel element (2 ty + sy, 2 tx + sx)
              float* dst = a.dw + ((int64_t)chn * a.Bc + b) * 16 + (2 * ty) * 4 + 2 * tx;
              if (v2ok) {
                atomicAdd(reinterpret_cast<float2*>(dst), make_float2(v[jb * 4 + 0], v[jb * 4 + 1]));
                atomicAdd(reinterpret_cast<float2*>(dst + 4), make_float2(v[jb * 4 + 2], v[jb * 4 + 3]));
              } else {
                atomicAdd(dst, v[jb * 4 + 0]); atomicAdd(dst + 1, v[jb * 4 + 1]);
                atomicAdd(dst + 4, v[jb * 4 + 2]); atomicAdd(dst + 5, v[jb * 4 + 3]);
              }
            }
          }
        }
      }
    } else {
    const int ch = a0 + tid;
    // The split-K partial sums go to dW with 16-byte vector atomics: one per (channel pair row) instead
    // of four scalar ones.  With scalar atomics the deep layers (few positions, 10^5 weights, 36 splits)
    // spent most of their ~110 us in this epilogue, whatever their size.
    const bool vec = (reinterpret_cast<uintptr_t>(a.dw) & 15) == 0;
#pragma unroll 1
    for (int g = 0; g < NT / 16; ++g) {
      if (DIM == 2) {
        // one big-map channel at a time (4 taps x 4 sub-positions = its 4x4 kernel): 16 live registers,
        // so the epilogue does not set the kernel's register count (3 CTAs per SM on the first layer)
#pragma unroll 1
        for (int jb = 0; jb < 4; ++jb) {
          const int b = (c0 + g * 16) / 4 + jb;
          if (b >= a.Bc) break;                             // warp-uniform
          float v[4][4];                                    // [tap][sub-position]
#pragma unroll
          for (int tap = 0; tap < 4; ++tap) tmem_ld4(trow + tap * NT + g * 16 + jb * 4, v[tap]);
          if (ch < a.A && nkb > 0) {
            float* dst = a.dw + ((int64_t)ch * a.Bc + b) * 16;
#pragma unroll
            for (int ky = 0; ky < 4; ++ky) {
              // kernel row ky = 2*ty + sy: taps (ty,0),(ty,1), sub-positions (sy,0),(sy,1)
              const int ty = ky >> 1, sy = ky & 1;
              const float4 r = make_float4(v[2 * ty][2 * sy], v[2 * ty][2 * sy + 1], v[2 * ty + 1][2 * sy], v[2 * ty + 1][2 * sy + 1]);
              if (vec) {
                atomicAdd(reinterpret_cast<float4*>(dst + ky * 4), r);
              } else {
                atomicAdd(dst + ky * 4 + 0, r.x); atomicAdd(dst + ky * 4 + 1, r.y);
                atomicAdd(dst + ky * 4 + 2, r.z); atomicAdd(dst + ky * 4 + 3, r.w);
              }
            }
          }
        }
      } else {
        float v[16];
        tmem_ld16(trow + g * 16, v);
        if (ch < a.A && nkb > 0) {
#pragma unroll
          for (int jb = 0; jb < 4; ++jb) {
            const int b = (c0 + g * 16) / 4 + jb;
            if (b < a.Bc) {
              float* dst = a.dw + ((int64_t)ch * a.Bc + b) * 4;
              if (vec) {
                atomicAdd(reinterpret_cast<float4*>(dst), make_float4(v[jb * 4], v[jb * 4 + 1], v[jb * 4 + 2], v[jb * 4 + 3]));
              } else {
#pragma unroll
                for (int t = 0; t < 4; ++t) atomicAdd(dst + t, v[jb * 4 + t]);
              }
            }
          }
        }
      }
    }
    }
    }
  } else if (FUSED && warp < MMAW) {
    // ------------------------------------------------ data-gradient epilogue (warps 4-7 = TMEM lane quarters 0-3)
    const int row = (warp & 3) * 32 + lane;
    const uint32_t trow = tmem + ((uint32_t)((warp & 3) * 32) << 16) + DCOL;
    for (int it = 0; it < nkb; ++it) {
      const uint32_t buf = (uint32_t)it & 1u;
      const uint32_t q = (uint32_t)((kb0 + it) * KP) + (uint32_t)row;
      bool ok = q < (uint32_t)a.Q;
      uint32_t n = 0, pos = 0;
      if (ok) {
        if (DIM == 2) {
          n = fdiv(q, a.d_pp);
          const uint32_t r = q - n * (uint32_t)(PH * PW);
          const uint32_t by = fdiv(r, a.d_pw), bx = r - by * (uint32_t)PW;
          ok = by < (uint32_t)a.h && bx < (uint32_t)a.w;
          pos = by * (uint32_t)a.w + bx;
        } else {
          n = fdiv(q, a.d_w);
          pos = q - n * (uint32_t)a.w;
        }
      }
      const int64_t chs = DIM == 2 ? (int64_t)a.h * a.w : (int64_t)a.w;   // channel stride of the small map
      const float* sp = a.small_ + (int64_t)n * a.small_ns + pos;
      float* op = a.dz + (int64_t)n * a.dz_ns + pos;
      mbar_wait(&dacc_full[buf], ((uint32_t)it >> 1) & 1u);
      fence_after();
#pragma unroll 1
      for (int h8 = 0; h8 < NTD / 8; ++h8) {
        if (h8 * 8 >= a.A) break;                    // warp-uniform
        float ax[8], v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) ax[e] = (ok && h8 * 8 + e < a.A) ? __ldg(sp + (int64_t)(h8 * 8 + e) * chs) : 0.f;
        tmem_ld8(trow + buf * NTD + h8 * 8, v);
#pragma unroll
        for (int e = 0; e < 8; ++e)
          if (ok && h8 * 8 + e < a.A) op[(int64_t)(h8 * 8 + e) * chs] = v[e] * delu_from_out(ax[e]);
      }
      fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&dacc_empty[buf]);
    }
  } else {
    // ------------------------------------------------ MMA issuer: warp-uniform loop, one elected lane issues
    // M = 64 when the layer has at most 16 output channels (rows 0-15 of the accumulator sit in TMEM lanes
    // 0-15 for either M): the M-side operand fetch from shared memory (4 KB per MMA at M = 128, of which
    // 8 rows are real in the first layer) is what bounds the wide first layers.
    const uint32_t idesc = make_idesc(NT, 1, 1, FOLD ? (32 * a.acols <= 64 ? 64 : 128) : (a.A <= 16 ? 64 : 128));
    const uint32_t leader = elect_one();
    const uint32_t idesc_d = make_idesc(NTD, 0, 0);
    if (FUSED) mbar_wait(&dimg_bar, 0);
    Ring ring{0, 0};
    for (int it = 0; it < nkb; ++it, ring.next(NS)) {
      const int s = ring.s;
      mbar_wait(&full_bar[s], ring.ph);
      fence_after();
      if (FUSED) {
        // data gradient first (its epilogue warps start while the weight-gradient MMAs of the block are issued): the Z
        // tile as a K-major operand, chunk columns ZS * 16 bytes apart (igemm_down.cu), against the resident image
        const uint32_t buf = (uint32_t)it & 1u;
        mbar_wait(&dacc_empty[buf], (((uint32_t)it >> 1) & 1u) ^ 1u);
        fence_after();
        const uint32_t zh = smem_u32(smem + (size_t)s * stage_bytes) + 2 * SBYTES;
        const uint32_t bh = smem_u32(smem + a.dimg_off);
        const uint64_t dah = make_desc(zh, ZS * 16, 128), dal = make_desc(zh + zbytes, ZS * 16, 128);
        const uint64_t dbh = make_desc(bh, NTD * 16, 128), dbl = make_desc(bh + DIMG / 2, NTD * 16, 128);
#pragma unroll 1
        for (int tap = 0; tap < TD; ++tap) {
          const uint32_t shift = DIM == 2 ? (uint32_t)((tap >> 1) * PW + (tap & 1)) : 0u;   // the tile carries a halo
#pragma unroll
          for (int ks = 0; ks < NT / 16; ++ks) {
            const uint32_t ao = (uint32_t)(2 * ks) * ZS + shift, bo = (uint32_t)(tap * 4 + 2 * ks) * NTD;
            mma_split3_warp(tmem + DCOL + buf * NTD, desc_off(dah, ao), desc_off(dal, ao), desc_off(dbh, bo), desc_off(dbl, bo),
                            idesc_d, (tap > 0 || ks > 0) ? 1u : 0u, leader);
          }
        }
        commit_warp(&dacc_full[buf], leader);
      }
      const uint32_t shi = smem_u32(smem + (size_t)s * stage_bytes);
      const uint32_t slo = shi + SBYTES;
      const uint32_t zhi = slo + SBYTES;
      const uint32_t zlo = zhi + zbytes;
      const uint64_t dsh = make_desc(shi, 128, KP * 16), dsl = make_desc(slo, 128, KP * 16);
      const uint64_t dzh = make_desc(zhi, 128, ZS * 16), dzl = make_desc(zlo, 128, ZS * 16);
      // rolled on purpose: keeps the issuing warp's code small (instruction-cache footprint)
#pragma unroll 1
      for (int tap = 0; tap < T; ++tap) {
        const uint32_t shift = DIM == 2 ? (uint32_t)((tap >> 1) * PW + (tap & 1)) : 0u;
        // the K steps are unrolled (offsets become immediates: ~6 instructions per MMA instead of 13; at
        // 96 MMAs per 128 positions the issuing warp is what bounds the first 2-D layer), the taps are not
#pragma unroll
        for (int ks = 0; ks < KP / 16; ++ks) {
          const uint32_t ao = (uint32_t)ks * 16;                   // 16-byte units
          const uint32_t bo = (uint32_t)ks * 16 + shift;
          mma_split3_warp(tmem + tap * NT, desc_off(dsh, ao), desc_off(dsl, ao), desc_off(dzh, bo), desc_off(dzl, bo), idesc,
                          (it > 0 || ks > 0) ? 1u : 0u, leader);
        }
      }
      commit_warp(&empty_bar[s], leader);
    }
    commit_warp(&acc_bar, leader);
  }
  fence_before();
  __syncthreads();
  if (warp == MMAW) tmem_dealloc(tmem, TMEM_COLS);
}

template <int DIM, int NT, int KP, bool PRE = false, bool FOLD = false, bool FUSED = false>
int launch_wgrad_t(WgArgs a, int64_t /*splits_hint*/, int mtiles, cudaStream_t st) {
  if (PRE) {
    const PlaneGeom pg = plane_geom(DIM, a.N, a.Bc, a.h, a.w);
    const uint8_t* base = reinterpret_cast<const uint8_t*>(a.big);
    if (int rc = make_plane_tmap(&a.tm_hi, base, pg.Qs, pg.chunks, a.zslots, NT / 8)) return rc;
    if (int rc = make_plane_tmap(&a.tm_lo, base + pg.half_bytes, pg.Qs, pg.chunks, a.zslots, NT / 8)) return rc;
  }
  const size_t stage = (size_t)2 * a.scols * KP * 16 + (size_t)2 * (NT / 8) * a.zslots * 16;
  // slack: the padding row groups of the last stage's S tiles are read (and ignored) up to 16 groups
  // (M = 64 instances - at most 16 small-map channels, or the folded 8-channel layer - walk 8 groups only)
  const bool m64 = FOLD ? 32 * a.acols <= 64 : a.A <= 16;
  const size_t reach = (size_t)a.scols * KP * 16 + (size_t)(m64 ? 8 : 16) * KP * 16;     // from the stage start
  size_t smem = stage * a.nstage + (reach > stage ? reach - stage : 0) + 256;
  if (FUSED) { a.dimg_off = (uint32_t)((smem + 127) / 128 * 128); smem = a.dimg_off + (size_t)2 * (DIM == 2 ? 4 : 1) * 4 * 16 * 16; }
  LSHM_CUDA(cudaFuncSetAttribute(igemm_wgrad_kernel<DIM, NT, KP, PRE, FOLD, FUSED>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "igemm_wgrad");
  // split-K so that the grid is ONE wave of resident CTAs (the first version assumed 3 CTAs per SM: where only 2 fit,
  // e.g. the 12-channel 1-D layer, 1.46 waves left a third of the run to a half-empty machine - ncu, r2_ncu_layer2.md)
  // resident CTAs per SM: the launch bound (registers) and the shared memory (227 KB per SM, 1 KB reserved per CTA).
  // (cudaOccupancyMaxActiveBlocksPerMultiprocessor answered 1 for these kernels on the driver of the test box.)
  const int occ_cache = (int)std::max<size_t>(1, std::min<size_t>((KP == 128 && !(FUSED && (DIM == 2 || !PRE))) ? 3 : 2, (size_t)(227 * 1024) / (smem + 1024)));
  const int64_t tiles = (int64_t)mtiles * a.ntiles;
  int64_t splits = std::max<int64_t>(1, ((int64_t)sm_count() * occ_cache) / tiles);
  splits = std::min(splits, std::max<int64_t>(1, a.kblocks / 4));   // at least 4 K blocks per CTA
  a.kb_per_cta = ceil_div(a.kblocks, splits);
  splits = ceil_div(a.kblocks, a.kb_per_cta);
  a.nstage = (int)std::min<int64_t>(a.nstage, std::max<int64_t>(1, a.kb_per_cta));
  dim3 grid((unsigned)splits, (unsigned)(mtiles * a.ntiles));
  igemm_wgrad_kernel<DIM, NT, KP, PRE, FOLD, FUSED><<<grid, wg_threads(DIM, FUSED, PRE), smem, st>>>(a);
  LSHM_CHECK_LAUNCH("igemm_wgrad");
  return LSHM_OK;
}

int launch_wgrad(int dim, WgArgs a, cudaStream_t st, bool planes = false, bool fused = false) {
  const int Kc = 4 * a.Bc;
  const int k16 = (Kc + 15) / 16 * 16;
  const int NT = k16 <= 16 ? 16 : (k16 <= 32 ? 32 : (k16 <= 48 ? 48 : 96));
  a.ntiles = (Kc + NT - 1) / NT;
  const int mtiles = (a.A + 127) / 128;
  a.Q = dim == 2 ? a.N * (int64_t)(a.h + 1) * (a.w + 1) : a.N * (int64_t)a.w;
  a.d_pp = make_fastdiv((uint32_t)((a.h + 1) * (a.w + 1))); a.d_pw = make_fastdiv((uint32_t)(a.w + 1)); a.d_w = make_fastdiv((uint32_t)a.w);
  LSHM_REQUIRE(a.Q < (1LL << 31) - 4096, "lshm_wgrad: too many positions (%lld) for one call; split the batch", (long long)a.Q);
  a.scols = std::min(16, (std::min(a.A, 128) + 7) / 8);
  a.vec_ok = ((reinterpret_cast<uintptr_t>(a.big) & 15) == 0 && (a.big_ns & 3) == 0) ? 1 : 0;
  const int KP = (a.scols <= 2 && NT <= 32) ? 128 : 64;
  a.kblocks = ceil_div(a.Q, KP);
  // taps folded into M (see the kernel): the 8 / 12-channel layers, whose un-folded form is bound by the tensor pipe
  static const bool nofold = getenv("LSHM_WGRAD_NOFOLD") != nullptr;          // experiment switch
  const bool fold = dim == 2 && a.A <= 8 && KP == 128 && !nofold;   // (A = 12: staging the small map 4x costs more than it saves)
  if (fold) { a.acols = (a.A + 7) / 8; a.scols = 4 * a.acols; }
  a.zslots = (dim == 2 && (!fold || fused)) ? (KP + a.w + 2 + 7) / 8 * 8 : KP;   // (fused: the data gradient reads a halo)
  if (planes) a.zslots = (a.zslots + PLANE_ROW - 1) / PLANE_ROW * PLANE_ROW;   // whole 512-byte rows of the tensor map
  a.d_zs = make_fastdiv((uint32_t)a.zslots);
  const size_t stage = (size_t)2 * a.scols * KP * 16 + (size_t)2 * (NT / 8) * a.zslots * 16;
  a.nstage = (int)std::min<size_t>(3, std::max<size_t>(2, (72 * 1024) / stage));
  const int64_t tiles = (int64_t)mtiles * a.ntiles;
  int64_t splits = std::max<int64_t>(1, ceil_div((int64_t)sm_count() * 3, tiles));
  splits = std::min(splits, std::max<int64_t>(1, a.kblocks / 4));   // at least 4 K blocks per CTA
  a.kb_per_cta = ceil_div(a.kblocks, splits);
  splits = ceil_div(a.kblocks, a.kb_per_cta);
  a.nstage = (int)std::min<int64_t>(a.nstage, std::max<int64_t>(1, a.kb_per_cta));
  if (planes) {
    LSHM_REQUIRE(NT <= 32 && KP == 128 && a.zslots <= 256, "lshm_wgrad*_planes: operand planes serve the first layers (A <= 16, Bc <= 8)");
    if (fused && dim == 2) {
      LSHM_REQUIRE(fold, "lshm_tconv_bwd2d_planes: at most 8 small-map channels");
      if (NT == 16) return launch_wgrad_t<2, 16, 128, true, true, true>(a, splits, mtiles, st);
      return launch_wgrad_t<2, 32, 128, true, true, true>(a, splits, mtiles, st);
    }
    if (dim == 2 && fold) { if (NT == 16) return launch_wgrad_t<2, 16, 128, true, true>(a, splits, mtiles, st); return launch_wgrad_t<2, 32, 128, true, true>(a, splits, mtiles, st); }
    if (dim == 2) { if (NT == 16) return launch_wgrad_t<2, 16, 128, true>(a, splits, mtiles, st); return launch_wgrad_t<2, 32, 128, true>(a, splits, mtiles, st); }
    if (fused) {
      LSHM_REQUIRE(a.A <= 16, "lshm_tconv_bwd1d_planes: at most 16 small-map channels");
      if (NT == 16) return launch_wgrad_t<1, 16, 128, true, false, true>(a, splits, mtiles, st);
      return launch_wgrad_t<1, 32, 128, true, false, true>(a, splits, mtiles, st);
    }
    if (NT == 16) return launch_wgrad_t<1, 16, 128, true>(a, splits, mtiles, st);
    return launch_wgrad_t<1, 32, 128, true>(a, splits, mtiles, st);
  }
  if (fused) {
    // fp32 big map: the second layers (12 -> 8 channels); the kernel then gathers the big-map tile once for both products
    LSHM_REQUIRE(!fold && KP == 128 && NT == 32 && a.A <= 16, "lshm_tconv_bwd*: 9..16 small-map channels and 8 big-map channels");
    if (dim == 2) return launch_wgrad_t<2, 32, 128, false, false, true>(a, splits, mtiles, st);
    return launch_wgrad_t<1, 32, 128, false, false, true>(a, splits, mtiles, st);
  }
  if (fold) {
    if (NT == 16) return launch_wgrad_t<2, 16, 128, false, true>(a, splits, mtiles, st);
    return launch_wgrad_t<2, 32, 128, false, true>(a, splits, mtiles, st);
  }
#define LW(D, NTV, KPV) return launch_wgrad_t<D, NTV, KPV>(a, splits, mtiles, st)
  if (dim == 2) {
    switch (NT) {
      case 16: if (KP == 128) LW(2, 16, 128); else LW(2, 16, 64);
      case 32: if (KP == 128) LW(2, 32, 128); else LW(2, 32, 64);
      case 48: LW(2, 48, 64);
      default: LW(2, 96, 64);
    }
  } else {
    switch (NT) {
      case 16: if (KP == 128) LW(1, 16, 128); else LW(1, 16, 64);
      case 32: if (KP == 128) LW(1, 32, 128); else LW(1, 32, 64);
      case 48: LW(1, 48, 64);
      default: LW(1, 96, 64);
    }
  }
#undef LW
}

}  // namespace
}  // namespace lshm

using namespace lshm;

extern "C" {

int lshm_wgrad2d(const float* small_, int64_t small_ns, const float* big, int64_t big_ns,
                 float* dw, int64_t N, int A, int Bc, int h, int w_, lshm_stream_t stream) {
  LSHM_REQUIRE(small_ && big && dw, "lshm_wgrad2d: null pointer");
  LSHM_REQUIRE(N >= 0 && A > 0 && Bc > 0 && (Bc & 3) == 0 && h > 0 && w_ > 0 && w_ <= 512, "lshm_wgrad2d: bad sizes");
  cudaStream_t st = as_stream(stream);
  LSHM_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)A * Bc * 16, st), "lshm_wgrad2d");
  if (N == 0) return LSHM_OK;
  WgArgs a{};
  a.small_ = small_; a.small_ns = small_ns; a.big = big; a.big_ns = big_ns; a.dw = dw;
  a.N = N; a.A = A; a.Bc = Bc; a.h = h; a.w = w_; a.pad = 0;
  return launch_wgrad(2, a, st);
}

int lshm_wgrad1d(const float* small_, int64_t small_ns, const float* big, int64_t big_ns,
                 float* dw, int64_t N, int A, int Bc, int l, int pad, lshm_stream_t stream) {
  LSHM_REQUIRE(small_ && big && dw, "lshm_wgrad1d: null pointer");
  LSHM_REQUIRE(N >= 0 && A > 0 && Bc > 0 && (Bc & 3) == 0 && l > 0 && (pad == 0 || pad == 1), "lshm_wgrad1d: bad sizes");
  cudaStream_t st = as_stream(stream);
  LSHM_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)A * Bc * 4, st), "lshm_wgrad1d");
  if (N == 0) return LSHM_OK;
  WgArgs a{};
  a.small_ = small_; a.small_ns = small_ns; a.big = big; a.big_ns = big_ns; a.dw = dw;
  a.N = N; a.A = A; a.Bc = Bc; a.h = 1; a.w = l; a.pad = pad;
  return launch_wgrad(1, a, st);
}

// Same kernels with the big map given as operand planes (see lshm_down*_planes).
int lshm_wgrad2d_planes(const float* small_, int64_t small_ns, const void* planes,
                        float* dw, int64_t N, int A, int Bc, int h, int w_, lshm_stream_t stream) {
  LSHM_REQUIRE(small_ && planes && dw, "lshm_wgrad2d_planes: null pointer");
  LSHM_REQUIRE(N >= 0 && A > 0 && Bc > 0 && (Bc & 3) == 0 && h > 0 && w_ > 0 && w_ <= 118, "lshm_wgrad2d_planes: bad sizes");
  cudaStream_t st = as_stream(stream);
  LSHM_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)A * Bc * 16, st), "lshm_wgrad2d_planes");
  if (N == 0) return LSHM_OK;
  WgArgs a{};
  a.small_ = small_; a.small_ns = small_ns; a.big = reinterpret_cast<const float*>(planes); a.big_ns = 0; a.dw = dw;
  a.N = N; a.A = A; a.Bc = Bc; a.h = h; a.w = w_; a.pad = 0;
  return launch_wgrad(2, a, st, true);
}

int lshm_wgrad1d_planes(const float* small_, int64_t small_ns, const void* planes,
                        float* dw, int64_t N, int A, int Bc, int l, lshm_stream_t stream) {
  LSHM_REQUIRE(small_ && planes && dw, "lshm_wgrad1d_planes: null pointer");
  LSHM_REQUIRE(N >= 0 && A > 0 && Bc > 0 && (Bc & 3) == 0 && l > 0, "lshm_wgrad1d_planes: bad sizes");
  cudaStream_t st = as_stream(stream);
  LSHM_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)A * Bc * 4, st), "lshm_wgrad1d_planes");
  if (N == 0) return LSHM_OK;
  WgArgs a{};
  a.small_ = small_; a.small_ns = small_ns; a.big = reinterpret_cast<const float*>(planes); a.big_ns = 0; a.dw = dw;
  a.N = N; a.A = A; a.Bc = Bc; a.h = 1; a.w = l; a.pad = 0;
  return launch_wgrad(1, a, st, true);
}

// Weight gradient AND data gradient of a k4/s4 transposed conv whose output gradient is given as operand planes:
// = lshm_wgrad1d_planes(small_, planes -> dw) + lshm_down1d_planes(planes, wimg_down, aux = small_, LSHM_EPI_DELU -> dz)
// with the planes read once.
int lshm_tconv_bwd1d_planes(const float* small_, int64_t small_ns, const void* planes, const void* wimg_down,
                            float* dz, int64_t dz_ns, float* dw, int64_t N, int A, int Bc, int l, lshm_stream_t stream) {
  LSHM_REQUIRE(small_ && planes && wimg_down && dz && dw, "lshm_tconv_bwd1d_planes: null pointer");
  LSHM_REQUIRE(N >= 0 && A > 0 && A <= 16 && Bc > 0 && (Bc & 3) == 0 && Bc <= 8 && l > 0, "lshm_tconv_bwd1d_planes: bad sizes (A <= 16, Bc in {4, 8})");
  LSHM_REQUIRE((reinterpret_cast<uintptr_t>(wimg_down) & 15) == 0, "lshm_tconv_bwd1d_planes: weight image must be 16-byte aligned");
  cudaStream_t st = as_stream(stream);
  LSHM_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)A * Bc * 4, st), "lshm_tconv_bwd1d_planes");
  if (N == 0) return LSHM_OK;
  WgArgs a{};
  a.small_ = small_; a.small_ns = small_ns; a.big = reinterpret_cast<const float*>(planes); a.big_ns = 0; a.dw = dw;
  a.dimg = reinterpret_cast<const uint8_t*>(wimg_down); a.dz = dz; a.dz_ns = dz_ns;
  a.N = N; a.A = A; a.Bc = Bc; a.h = 1; a.w = l; a.pad = 0;
  return launch_wgrad(1, a, st, true, true);
}

// The same with the gradient w.r.t. the layer's output given as an fp32 map (the second-to-last transposed convs):
// = lshm_wgrad*d(small_, big -> dw) + lshm_down*d(big, wimg_down, aux = small_, LSHM_EPI_DELU -> dz), the big map
// gathered and converted once.
int lshm_tconv_bwd1d(const float* small_, int64_t small_ns, const float* big, int64_t big_ns, const void* wimg_down,
                     float* dz, int64_t dz_ns, float* dw, int64_t N, int A, int Bc, int l, lshm_stream_t stream) {
  LSHM_REQUIRE(small_ && big && wimg_down && dz && dw, "lshm_tconv_bwd1d: null pointer");
  LSHM_REQUIRE(N >= 0 && A > 8 && A <= 16 && Bc == 8 && l > 0, "lshm_tconv_bwd1d: bad sizes (9 <= A <= 16, Bc = 8)");
  LSHM_REQUIRE((reinterpret_cast<uintptr_t>(wimg_down) & 15) == 0, "lshm_tconv_bwd1d: weight image must be 16-byte aligned");
  cudaStream_t st = as_stream(stream);
  LSHM_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)A * Bc * 4, st), "lshm_tconv_bwd1d");
  if (N == 0) return LSHM_OK;
  WgArgs a{};
  a.small_ = small_; a.small_ns = small_ns; a.big = big; a.big_ns = big_ns; a.dw = dw;
  a.dimg = reinterpret_cast<const uint8_t*>(wimg_down); a.dz = dz; a.dz_ns = dz_ns;
  a.N = N; a.A = A; a.Bc = Bc; a.h = 1; a.w = l; a.pad = 0;
  return launch_wgrad(1, a, st, false, true);
}

int lshm_tconv_bwd2d(const float* small_, int64_t small_ns, const float* big, int64_t big_ns, const void* wimg_down,
                     float* dz, int64_t dz_ns, float* dw, int64_t N, int A, int Bc, int h, int w_, lshm_stream_t stream) {
  LSHM_REQUIRE(small_ && big && wimg_down && dz && dw, "lshm_tconv_bwd2d: null pointer");
  LSHM_REQUIRE(N >= 0 && A > 8 && A <= 16 && Bc == 8 && h > 0 && w_ > 0 && w_ <= 126, "lshm_tconv_bwd2d: bad sizes (9 <= A <= 16, Bc = 8, w <= 126)");
  LSHM_REQUIRE((reinterpret_cast<uintptr_t>(wimg_down) & 15) == 0, "lshm_tconv_bwd2d: weight image must be 16-byte aligned");
  cudaStream_t st = as_stream(stream);
  LSHM_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)A * Bc * 16, st), "lshm_tconv_bwd2d");
  if (N == 0) return LSHM_OK;
  WgArgs a{};
  a.small_ = small_; a.small_ns = small_ns; a.big = big; a.big_ns = big_ns; a.dw = dw;
  a.dimg = reinterpret_cast<const uint8_t*>(wimg_down); a.dz = dz; a.dz_ns = dz_ns;
  a.N = N; a.A = A; a.Bc = Bc; a.h = h; a.w = w_; a.pad = 0;
  return launch_wgrad(2, a, st, false, true);
}

int lshm_tconv_bwd2d_planes(const float* small_, int64_t small_ns, const void* planes, const void* wimg_down,
                            float* dz, int64_t dz_ns, float* dw, int64_t N, int A, int Bc, int h, int w_, lshm_stream_t stream) {
  LSHM_REQUIRE(small_ && planes && wimg_down && dz && dw, "lshm_tconv_bwd2d_planes: null pointer");
  LSHM_REQUIRE(N >= 0 && A > 0 && A <= 8 && Bc > 0 && (Bc & 3) == 0 && Bc <= 8 && h > 0 && w_ > 0 && w_ <= 118,
               "lshm_tconv_bwd2d_planes: bad sizes (A <= 8, Bc in {4, 8}, w <= 118)");
  LSHM_REQUIRE((reinterpret_cast<uintptr_t>(wimg_down) & 15) == 0, "lshm_tconv_bwd2d_planes: weight image must be 16-byte aligned");
  cudaStream_t st = as_stream(stream);
  LSHM_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)A * Bc * 16, st), "lshm_tconv_bwd2d_planes");
  if (N == 0) return LSHM_OK;
  WgArgs a{};
  a.small_ = small_; a.small_ns = small_ns; a.big = reinterpret_cast<const float*>(planes); a.big_ns = 0; a.dw = dw;
  a.dimg = reinterpret_cast<const uint8_t*>(wimg_down); a.dz = dz; a.dz_ns = dz_ns;
  a.N = N; a.A = A; a.Bc = Bc; a.h = h; a.w = w_; a.pad = 0;
  return launch_wgrad(2, a, st, true, true);
}

}  // extern "C"
