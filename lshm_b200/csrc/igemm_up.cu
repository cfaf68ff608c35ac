// "up" implicit GEMM on tcgen05: ConvTranspose2d(k4,s2,p1) / ConvTranspose1d(k4,s4,p0) forward and the
// dgrad of Conv2d(k4,s2,p1) / Conv1d(k4,s4,p1).  Reference semantics: F.conv_transpose2d /
// F.conv_transpose1d at /root/reference/src/lofar_models.py:93-98,:178-183 (and the autograd of
// :73-78,:158-163).
//
// 2-D: an output pixel (2m+ry, 2n+rx) receives exactly 2x2 taps, so the four parity classes are
// four GEMMs over the SAME staged tile of the small map S (positions of the zero-padded
// (h+1)x(w+1) grid, rows at a 16-byte pitch, channels = K):
//     out[(2m+ry,2n+rx), b] = sum_{d,e} sum_a S[q + dy(ry,d)*(w+1) + dx(rx,e), a] * W[a,b,ky(ry,d),kx(rx,e)]
// with (ry=0: ky=1,dy=0 | ky=3,dy=-1) (ry=1: ky=0,dy=+1 | ky=2,dy=0).  The nine shifted views of S
// are nine descriptor start addresses into one shared-memory copy; the four classes accumulate in
// four TMEM column ranges and the epilogue writes complete 2x2 output blocks (float2 stores).
// 1-D (k=s=4): out[b, 4i+t-pad] = sum_a S[i,a] W[a,b,t], a plain GEMM with N = (b,t).
#include "conv_geom.cuh"

namespace lshm {
namespace {

using namespace tc;

struct UpArgs {
  const float* small_; int64_t small_ns;
  const uint8_t* wimg;
  const float* bias;
  const float* aux; int64_t aux_ns;
  float* big; int64_t big_ns;
  int64_t N; int A; int Bc; int h; int w; int pad; int epi;
  int slots; int nstage; int64_t Q; int64_t mtiles; int ntn;
  FastDiv d_pp, d_pw, d_w, d_ntn;   // divisors (h+1)(w+1), w+1, w, ntn
};

// warps 0-3 epilogue, 4-7 producers, 8 MMA issuer, 9 weight loader, (2-D only) 10 second MMA issuer
// G = 2 adds a second group of four producer warps after those (alternate stages, see igemm_down.cu)
constexpr int up_threads(int dim, int g = 1) { return (dim == 2 ? 352 : 320) + 128 * (g - 1); }
constexpr int UP_MAXST = 6;

template <int EPI>
__device__ __forceinline__ float epi_apply(float acc, float bias, float aux) {
  float r = acc + bias;
  if (EPI == LSHM_EPI_ELU) r = elu_fast(r);
  else if (EPI == LSHM_EPI_DELU) r *= delu_from_out(aux);
  return r;
}

// One accumulator set -> bias / activation -> global, for the thread's position (n, m, x).  The
// epilogue mode is a template parameter (dispatched once per tile by a warp-uniform switch): with a
// run-time mode every element carried the predicated-off instructions of the other two modes.
template <int DIM, int NT, int EPI>
__device__ __forceinline__ void up_epilogue_tile(const UpArgs& a, uint32_t trow, int nt, int64_t n, int m, int x, bool ok) {
  if (DIM == 2) {
    const int W = 2 * a.w;
    const int64_t HW = 4 * (int64_t)a.h * a.w;
    float* outp = a.big + n * a.big_ns + (int64_t)(2 * m) * W + 2 * x;
    const float* auxp = EPI == LSHM_EPI_DELU ? a.aux + n * a.aux_ns + (int64_t)(2 * m) * W + 2 * x : nullptr;
    // eight channels (x four parity classes = 32 accumulator registers) at a time: the 16-wide version
    // held 64 and spilled at the kernel's register cap
#pragma unroll 1
    for (int h8 = 0; h8 < NT / 8; ++h8) {
      const int b0 = nt * NT + h8 * 8;
      const int nch = min(8, a.Bc - b0);
      if (nch <= 0) break;                       // warp-uniform
      float v[4][8];
#pragma unroll
      for (int c = 0; c < 4; ++c) tmem_ld8(trow + c * NT + h8 * 8, v[c]);
      if (ok) {
        float* op = outp + (int64_t)b0 * HW;
        const float* xp = EPI == LSHM_EPI_DELU ? auxp + (int64_t)b0 * HW : nullptr;
        const float* bp = a.bias != nullptr ? a.bias + b0 : nullptr;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (j < nch) {
            const float bs = bp != nullptr ? __ldg(bp + j) : 0.f;
            float2 ax0 = make_float2(0.f, 0.f), ax1 = make_float2(0.f, 0.f);
            if (EPI == LSHM_EPI_DELU) {
              ax0 = *reinterpret_cast<const float2*>(xp + (int64_t)j * HW);
              ax1 = *reinterpret_cast<const float2*>(xp + (int64_t)j * HW + W);
            }
            // class index = ry*2 + rx
            const float2 o0 = make_float2(epi_apply<EPI>(v[0][j], bs, ax0.x), epi_apply<EPI>(v[1][j], bs, ax0.y));
            const float2 o1 = make_float2(epi_apply<EPI>(v[2][j], bs, ax1.x), epi_apply<EPI>(v[3][j], bs, ax1.y));
            *reinterpret_cast<float2*>(op + (int64_t)j * HW) = o0;
            *reinterpret_cast<float2*>(op + (int64_t)j * HW + W) = o1;
          }
        }
      }
    }
  } else {
    const int64_t Lb = 4 * (int64_t)a.w;
    float* outp = a.big + n * a.big_ns + 4 * (int64_t)x - a.pad;
    const float* auxp = EPI == LSHM_EPI_DELU ? a.aux + n * a.aux_ns + 4 * (int64_t)x - a.pad : nullptr;
#pragma unroll 1
    for (int g = 0; g < NT / 16; ++g) {
      const int bb0 = (nt * NT + g * 16) / 4;
      if (bb0 >= a.Bc) break;                    // warp-uniform
      float v[16];
      tmem_ld16(trow + g * 16, v);
      if (ok) {
#pragma unroll
        for (int jb = 0; jb < 4; ++jb) {
          const int b = bb0 + jb;
          if (b < a.Bc) {
            const float bs = a.bias != nullptr ? __ldg(a.bias + b) : 0.f;
            float* o = outp + b * Lb;
            if (a.pad == 0 && EPI != LSHM_EPI_DELU) {
              float4 r;
              r.x = epi_apply<EPI>(v[jb * 4 + 0], bs, 0.f); r.y = epi_apply<EPI>(v[jb * 4 + 1], bs, 0.f);
              r.z = epi_apply<EPI>(v[jb * 4 + 2], bs, 0.f); r.w = epi_apply<EPI>(v[jb * 4 + 3], bs, 0.f);
              *reinterpret_cast<float4*>(o) = r;
            } else {
#pragma unroll
              for (int t = 0; t < 4; ++t) {
                if (a.pad == 1 && x == 0 && t == 0) continue;        // position -1 does not exist
                const float ax = EPI == LSHM_EPI_DELU ? auxp[b * Lb + t] : 0.f;
                o[t] = epi_apply<EPI>(v[jb * 4 + t], bs, ax);
              }
              if (a.pad == 1 && x == a.w - 1) {                       // last position: no tap reaches it
                const float ax = EPI == LSHM_EPI_DELU ? auxp[b * Lb + 4] : 0.f;
                o[4] = epi_apply<EPI>(0.f, bs, ax);
              }
            }
          }
        }
      }
    }
  }
}

// Persistent, warp-specialised (same skeleton as igemm_down.cu): producers run ahead through the
// stage ring, the MMA warp alternates between two TMEM accumulator sets (each = 4 parity classes in
// 2-D), the epilogue warps drain one set while the next is being computed.
template <int DIM, int NT, int KC, int G>
__global__ void __launch_bounds__(up_threads(DIM, G), (DIM == 1 && G == 1 ? 3 : 2)) igemm_up_kernel(UpArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t full_bar[UP_MAXST], empty_bar[UP_MAXST], acc_full[2], acc_empty[2];
  __shared__ uint32_t tmem_base;
  constexpr int NCLS = DIM == 2 ? 4 : 1;
  constexpr int COMBOS = DIM == 2 ? 16 : 1;       // class x tap weight tiles per K block
  constexpr int CC = KC / 8;
  constexpr int NSLOT = DIM == 2 ? 3 : 1;         // staged slots per producer thread (1-D tiles: 128 slots)
  constexpr uint32_t IMG = 2u * COMBOS * CC * NT * 16;
  constexpr uint32_t TSET = NCLS * NT;            // TMEM columns of one accumulator set
  constexpr uint32_t TCOLS = 2 * TSET;
  constexpr uint32_t TMEM_COLS = TCOLS <= 32 ? 32 : (TCOLS <= 64 ? 64 : (TCOLS <= 128 ? 128 : (TCOLS <= 256 ? 256 : 512)));
  const int SLOTS = a.slots, NS = a.nstage;
  const uint32_t zbytes = (uint32_t)CC * SLOTS * 16;
  const uint32_t stage_bytes = 2 * zbytes + IMG;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int Apad = (a.A + 15) / 16 * 16;
  const int KB = (Apad + KC - 1) / KC;
  const int PW = a.w + 1, PH = a.h + 1;
  const int halo = DIM == 2 ? PW + 1 : 0;           // slots in front of the tile
  const int64_t total = a.mtiles * a.ntn;

  if (warp == 8) tmem_alloc(&tmem_base, TMEM_COLS);
  if (tid == 0) {
    // two MMA-issuing warps (each a serial one-thread chain) share the (class, tap) pairs of a stage
    for (int s = 0; s < UP_MAXST; ++s) { mbar_init(&full_bar[s], 5); mbar_init(&empty_bar[s], DIM == 2 ? 2 : 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&acc_full[b], DIM == 2 ? 2 : 1); mbar_init(&acc_empty[b], 4); }
    mbar_init_fence();
  }
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem = tmem_base;

  constexpr int G1W = (DIM == 2 ? 352 : 320) / 32;   // first warp of producer group 1 (= up_threads(DIM) / 32)
  if ((warp >= 4 && warp < 8) || warp >= G1W) {
    // ------------------------------------------------ producers: stage S (hi/lo bf16), K = channels
    const int grp = warp >= G1W ? 1 : 0;
    const int ptid = grp == 0 ? tid - 128 : tid - G1W * 32;
    uint32_t unit = 0;                                  // (item, K block) counter: group g fills units u % G == g
    const int64_t hw = DIM == 2 ? (int64_t)a.h * a.w : (int64_t)a.w;
    Ring ring{0, 0};
    if constexpr (DIM == 1 && G == 1 && KC == 16) {
      // Software-pipelined producer (1-D, one group): the loads of the NEXT (item, K block) unit are issued before the
      // current one is converted, so two units' worth of global loads are in flight per CTA.  With one unit in flight
      // the first / second layers were bound by the load -> convert -> hand-off latency chain (~1600 cycles per tile
      // and CTA, "long scoreboard" the top stall in profiles/r2_ncu_layer2.md), not by bandwidth.
      auto describe = [&](int64_t item, const float*& sp, bool& sv) {
        const int64_t q = (int64_t)fdiv((uint32_t)item, a.d_ntn) * 128 + ptid;
        sv = ptid < SLOTS && q < a.Q;
        sp = a.small_;
        if (sv) {
          const uint32_t uq = (uint32_t)q;
          const uint32_t n = fdiv(uq, a.d_w);
          sp = a.small_ + (int64_t)n * a.small_ns + (uq - n * (uint32_t)a.w);
        }
      };
      auto fetch = [&](const float* sp, bool sv, int kb, float (&v)[CC][8]) {
        const int ccb = (min(KC, Apad - kb * KC)) >> 3;
#pragma unroll
        for (int cc = 0; cc < CC; ++cc) {
          const int a0 = kb * KC + cc * 8;
          const float* p = sp + (int64_t)a0 * hw;
          const bool on = cc < ccb && sv;
          if (on && a0 + 8 <= a.A) {
#pragma unroll
            for (int e = 0; e < 8; ++e) { v[cc][e] = __ldg(p); p += hw; }
          } else {
#pragma unroll
            for (int e = 0; e < 8; ++e) { v[cc][e] = (on && a0 + e < a.A) ? __ldg(p) : 0.f; p += hw; }
          }
        }
      };
      int64_t item = blockIdx.x;
      int kb = 0;
      const float* sp = a.small_; bool sv = false;
      float v[CC][8];
      if (item < total) { describe(item, sp, sv); fetch(sp, sv, 0, v); }
      while (item < total) {
        int64_t nitem = item; int nkb = kb + 1;
        if (nkb == KB) { nkb = 0; nitem = item + gridDim.x; }
        const float* nsp = sp; bool nsv = sv;
        float vn[CC][8];
        if (nitem < total) {
          if (nitem != item) describe(nitem, nsp, nsv);
          fetch(nsp, nsv, nkb, vn);
        }
        const int s = ring.s;
        mbar_wait(&empty_bar[s], ring.ph ^ 1);
        uint8_t* zhi = smem + (size_t)s * stage_bytes;
        uint8_t* zlo = zhi + zbytes;
        const int ccb = (min(KC, Apad - kb * KC)) >> 3;
        if (ptid < SLOTS) {
#pragma unroll
          for (int cc = 0; cc < CC; ++cc) {
            if (cc < ccb) {
              uint4 hi, lo;
              split8(v[cc], hi, lo);
              *reinterpret_cast<uint4*>(zhi + ((size_t)cc * SLOTS + ptid) * 16) = hi;
              *reinterpret_cast<uint4*>(zlo + ((size_t)cc * SLOTS + ptid) * 16) = lo;
            }
          }
        }
        fence_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&full_bar[s]);
        ring.next(NS);
#pragma unroll
        for (int cc = 0; cc < CC; ++cc)
#pragma unroll
          for (int e = 0; e < 8; ++e) v[cc][e] = vn[cc][e];
        item = nitem; kb = nkb; sp = nsp; sv = nsv;
      }
    } else
    for (int64_t item = blockIdx.x; item < total; item += gridDim.x) {
      const int64_t q0 = (int64_t)fdiv((uint32_t)item, a.d_ntn) * 128;
      const float* sp[NSLOT]; bool sv[NSLOT];
#pragma unroll
      for (int i = 0; i < NSLOT; ++i) {
        const int s = ptid + i * 128;
        const int64_t q = q0 - halo + s;
        sv[i] = s < SLOTS && q >= 0 && q < a.Q;
        sp[i] = a.small_;
        if (sv[i]) {
          const uint32_t uq = (uint32_t)q;              // Q < 2^31 (launcher): 32-bit divisions only
          if (DIM == 2) {
            const uint32_t n = fdiv(uq, a.d_pp);
            const uint32_t r = uq - n * (uint32_t)(PH * PW);
            const int m = (int)fdiv(r, a.d_pw), x = (int)r - m * PW;
            sv[i] = m < a.h && x < a.w;
            sp[i] = a.small_ + (int64_t)n * a.small_ns + (int64_t)m * a.w + x;
          } else {
            const uint32_t n = fdiv(uq, a.d_w);
            sp[i] = a.small_ + (int64_t)n * a.small_ns + (uq - n * (uint32_t)a.w);
          }
        }
      }
      for (int kb = 0; kb < KB; ++kb, ring.next(NS), ++unit) {
        if (G == 2 && (int)(unit & 1) != grp) continue;   // the other group's stage
        const int s = ring.s;
        mbar_wait(&empty_bar[s], ring.ph ^ 1);
        uint8_t* zhi = smem + (size_t)s * stage_bytes;
        uint8_t* zlo = zhi + zbytes;
        const int ccb = (min(KC, Apad - kb * KC)) >> 3;
#pragma unroll
        for (int i = 0; i < NSLOT; ++i) {
          const int slot = ptid + i * 128;
          if (slot >= SLOTS) continue;
          float v[CC][8];
#pragma unroll
          for (int cc = 0; cc < CC; ++cc) {
            const int a0 = kb * KC + cc * 8;
            const float* p = sp[i] + (int64_t)a0 * hw;
            const bool on = cc < ccb && sv[i];
            if (on && a0 + 8 <= a.A) {
#pragma unroll
              for (int e = 0; e < 8; ++e) { v[cc][e] = __ldg(p); p += hw; }
            } else {
#pragma unroll
              for (int e = 0; e < 8; ++e) { v[cc][e] = (on && a0 + e < a.A) ? __ldg(p) : 0.f; p += hw; }
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
    }
  } else if (warp < 4) {
    // ------------------------------------------------ epilogue
    uint32_t tc_ = 0;
    for (int64_t item = blockIdx.x; item < total; item += gridDim.x, ++tc_) {
      const uint32_t mt = fdiv((uint32_t)item, a.d_ntn);
      const int nt = (int)((uint32_t)item - mt * (uint32_t)a.ntn);
      const int64_t q = (int64_t)mt * 128 + tid;
      bool ok = q < a.Q;
      int64_t n = 0; int m = 0, x = 0;
      if (ok) {
        const uint32_t uq = (uint32_t)q;
        if (DIM == 2) {
          const uint32_t un = fdiv(uq, a.d_pp);
          const uint32_t r = uq - un * (uint32_t)(PH * PW);
          m = (int)fdiv(r, a.d_pw); x = (int)r - m * PW;
          ok = m < a.h && x < a.w;
          n = un;
        } else {
          const uint32_t un = fdiv(uq, a.d_w);
          n = un; x = (int)(uq - un * (uint32_t)a.w);
        }
      }
      const uint32_t buf = tc_ & 1;
      mbar_wait(&acc_full[buf], (tc_ >> 1) & 1);
      fence_after();
      const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16) + buf * TSET;
      if (a.epi == LSHM_EPI_ELU) up_epilogue_tile<DIM, NT, LSHM_EPI_ELU>(a, trow, nt, n, m, x, ok);
      else if (a.epi == LSHM_EPI_DELU) up_epilogue_tile<DIM, NT, LSHM_EPI_DELU>(a, trow, nt, n, m, x, ok);
      else up_epilogue_tile<DIM, NT, LSHM_EPI_NONE>(a, trow, nt, n, m, x, ok);
      fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[buf]);
    }
  } else if (warp == 8 || warp == 10) {
    // ------------------------------------------------ MMA issuers.  The whole warp runs the (rolled)
    // loop so every descriptor lives in uniform registers and lane 0 only predicates the tcgen05
    // instructions: an unrolled single-lane version was 3800 SASS lines of waterfall loops that
    // evicted the other roles' code from the instruction cache (33% of stall samples were
    // "no instruction", profiles/r1_ncu_up2d_conv0.md).
    const int mw = warp == 8 ? 0 : 1;               // 2-D: issuer 0 takes classes 0,1, issuer 1 classes 2,3
    const uint32_t leader = elect_one();
    const uint32_t idesc = make_idesc(NT, 0, 0);
    const int cb0 = DIM == 2 ? mw * (COMBOS / 2) : 0, cb1 = DIM == 2 ? cb0 + COMBOS / 2 : COMBOS;
    uint32_t tc_ = 0;
    Ring ring{0, 0};
    for (int64_t item = blockIdx.x; item < total; item += gridDim.x, ++tc_) {
      const uint32_t buf = tc_ & 1;
      mbar_wait(&acc_empty[buf], ((tc_ >> 1) & 1) ^ 1);
      fence_after();
      const uint32_t tset = tmem + buf * TSET;
      for (int kb = 0; kb < KB; ++kb, ring.next(NS)) {
        const int s = ring.s;
        mbar_wait(&full_bar[s], ring.ph);
        fence_after();
        const uint32_t zhi = smem_u32(smem + (size_t)s * stage_bytes);
        const uint64_t dah = make_desc(zhi, SLOTS * 16, 128), dal = make_desc(zhi + zbytes, SLOTS * 16, 128);
        const uint64_t dbh = make_desc(zhi + 2 * zbytes, NT * 16, 128), dbl = make_desc(zhi + 2 * zbytes + IMG / 2, NT * 16, 128);
        const int ksteps = (min(KC, Apad - kb * KC)) >> 4;
#pragma unroll 1
        for (int cb = cb0; cb < cb1; ++cb) {
          // class = (ry, rx), tap = (d, e): row shift ry - d, column shift rx - e (header comment)
          const int cls = cb >> 2, tap = cb & 3;
          const uint32_t abase = DIM == 2 ? (uint32_t)(halo + ((cls >> 1) - (tap >> 1)) * PW + ((cls & 1) - (tap & 1))) : 0u;
          const uint32_t bbase = DIM == 2 ? (uint32_t)(cb * CC * NT) : 0u;
          const uint32_t td = DIM == 2 ? tset + cls * NT : tset;
          const bool first = DIM == 2 ? tap == 0 : true;                       // first tap of a class
#pragma unroll 1
          for (int ks = 0; ks < ksteps; ++ks) {
            const uint32_t ao = abase + (uint32_t)(2 * ks) * SLOTS, bo = bbase + (uint32_t)(2 * ks) * NT;
            mma_split3_warp(td, desc_off(dah, ao), desc_off(dal, ao), desc_off(dbh, bo), desc_off(dbl, bo), idesc,
                            (kb > 0 || !first || ks > 0) ? 1u : 0u, leader);
          }
        }
        commit_warp(&empty_bar[s], leader);
      }
      commit_warp(&acc_full[buf], leader);
    }
  } else {
    if (lane == 0) {
      Ring ring{0, 0};
      for (int64_t item = blockIdx.x; item < total; item += gridDim.x) {
        const uint32_t mt = fdiv((uint32_t)item, a.d_ntn);
        const int nt = (int)((uint32_t)item - mt * (uint32_t)a.ntn);
        for (int kb = 0; kb < KB; ++kb, ring.next(NS)) {
          const int s = ring.s;
          mbar_wait(&empty_bar[s], ring.ph ^ 1);
          mbar_arrive_expect_tx(&full_bar[s], IMG);
          bulk_g2s(smem + (size_t)s * stage_bytes + 2 * zbytes, a.wimg + ((size_t)nt * KB + kb) * IMG, IMG, &full_bar[s]);
        }
      }
    }
  }
  fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tmem, TMEM_COLS);
}

template <int DIM, int NT, int KC, int G>
int launch_up_t(UpArgs a, const UpGeom& g, cudaStream_t st) {
  const size_t stage = (size_t)2 * (KC / 8) * a.slots * 16 + g.img;
  const int64_t units = a.mtiles * g.ntiles * g.KB;
  constexpr int tcols = 2 * (DIM == 2 ? 4 : 1) * NT;
  const bool two = tcols <= 256;                       // TMEM allows two CTAs per SM
  int per_sm = 1;
  int ns = 0;
  if (DIM == 1 && G == 1 && tcols <= 128) { ns = (int)((74 * 1024) / stage); if (ns >= 2) per_sm = 3; }
  if (per_sm == 1 && two) { ns = (int)((110 * 1024) / stage); if (ns >= 2) per_sm = 2; }
  bool pair = per_sm > 1;
  if (!pair) ns = (int)((200 * 1024) / stage);
  ns = std::min(ns, UP_MAXST);
  LSHM_REQUIRE(ns >= 1, "igemm_up: tile does not fit in shared memory");
  a.nstage = (int)std::min<int64_t>(ns, std::max<int64_t>(1, units));
  const size_t smem = stage * a.nstage;
  LSHM_CUDA(cudaFuncSetAttribute(igemm_up_kernel<DIM, NT, KC, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "igemm_up");
  const int64_t grid = std::min<int64_t>(a.mtiles * g.ntiles, (int64_t)sm_count() * per_sm);
  igemm_up_kernel<DIM, NT, KC, G><<<(unsigned)grid, up_threads(DIM, G), smem, st>>>(a);
  LSHM_CHECK_LAUNCH("igemm_up");
  return LSHM_OK;
}

int launch_up(int dim, UpArgs a, cudaStream_t st) {
  const UpGeom g = up_geom(dim, a.A, a.Bc);
  a.slots = dim == 2 ? (128 + 2 * (a.w + 2) + 7) / 8 * 8 : 128;
  a.Q = dim == 2 ? a.N * (int64_t)(a.h + 1) * (a.w + 1) : a.N * (int64_t)a.w;
  a.d_pp = make_fastdiv((uint32_t)((a.h + 1) * (a.w + 1))); a.d_pw = make_fastdiv((uint32_t)(a.w + 1)); a.d_w = make_fastdiv((uint32_t)a.w);
  LSHM_REQUIRE(a.Q < (1LL << 31) - 4096, "lshm_up: too many positions (%lld) for one call; split the batch", (long long)a.Q);
  a.mtiles = ceil_div(a.Q, 128);
  a.ntn = g.ntiles;
  a.d_ntn = make_fastdiv((uint32_t)a.ntn);
  // two producer groups for the multi-K-block (deep) 1-D layers; in 2-D the 480-thread CTA would cap the
  // kernel at 64 registers (it needs 80) and measured slower
  const bool two = dim == 1 && g.KB >= 2;
#define LU(D, NTV, KCV) do { if (two) return launch_up_t<D, NTV, KCV, 2>(a, g, st); return launch_up_t<D, NTV, KCV, 1>(a, g, st); } while (0)
  if (dim == 2) {
    switch (g.NT) { case 16: LU(2, 16, 16); case 32: LU(2, 32, 16); default: LU(2, 48, 16); }
  } else {
    if (g.KC == 16) {       // <= 16 small-map channels (the first two layers): one K block of 16
      switch (g.NT) { case 16: return launch_up_t<1, 16, 16, 1>(a, g, st); case 32: return launch_up_t<1, 32, 16, 1>(a, g, st);
                      case 48: return launch_up_t<1, 48, 16, 1>(a, g, st); default: return launch_up_t<1, 96, 16, 1>(a, g, st); }
    }
    switch (g.NT) { case 16: LU(1, 16, 32); case 32: LU(1, 32, 32); case 48: LU(1, 48, 32); default: LU(1, 96, 32); }
  }
#undef LU
}

}  // namespace

}  // namespace lshm

using namespace lshm;

extern "C" {

int lshm_up2d(const float* small_, int64_t small_ns, const void* wimg, const float* bias,
              const float* aux, int64_t aux_ns, float* big, int64_t big_ns,
              int64_t N, int A, int Bc, int h, int w_, int epilogue, lshm_stream_t stream) {
  LSHM_REQUIRE(small_ && wimg && big, "lshm_up2d: null pointer");
  LSHM_REQUIRE(N >= 0 && A > 0 && Bc > 0 && h > 0 && w_ > 0, "lshm_up2d: bad sizes");
  // the producers stage 3 x 128 slots per tile (tile + two halo rows): 128 + 2 (w + 2) <= 384
  LSHM_REQUIRE(w_ <= 126, "lshm_up2d: small-map width %d too large (max 126)", w_);
  LSHM_REQUIRE(epilogue >= 0 && epilogue <= 2, "lshm_up2d: bad epilogue %d", epilogue);
  LSHM_REQUIRE(epilogue != LSHM_EPI_DELU || aux != nullptr, "lshm_up2d: DELU epilogue needs aux");
  LSHM_REQUIRE((reinterpret_cast<uintptr_t>(wimg) & 15) == 0 && (reinterpret_cast<uintptr_t>(big) & 7) == 0 &&
               (big_ns & 1) == 0 && (reinterpret_cast<uintptr_t>(aux) & 7) == 0 && (aux_ns & 1) == 0,
               "lshm_up2d: misaligned buffer");
  if (N == 0) return LSHM_OK;
  UpArgs a{};
  a.small_ = small_; a.small_ns = small_ns; a.wimg = reinterpret_cast<const uint8_t*>(wimg); a.bias = bias;
  a.aux = epilogue == LSHM_EPI_DELU ? aux : nullptr; a.aux_ns = aux_ns; a.big = big; a.big_ns = big_ns;
  a.N = N; a.A = A; a.Bc = Bc; a.h = h; a.w = w_; a.pad = 0; a.epi = epilogue;
  return launch_up(2, a, as_stream(stream));
}

int lshm_up1d(const float* small_, int64_t small_ns, const void* wimg, const float* bias,
              const float* aux, int64_t aux_ns, float* big, int64_t big_ns,
              int64_t N, int A, int Bc, int l, int pad, int epilogue, lshm_stream_t stream) {
  LSHM_REQUIRE(small_ && wimg && big, "lshm_up1d: null pointer");
  LSHM_REQUIRE(N >= 0 && A > 0 && Bc > 0 && l > 0 && (pad == 0 || pad == 1), "lshm_up1d: bad sizes");
  LSHM_REQUIRE(epilogue >= 0 && epilogue <= 2, "lshm_up1d: bad epilogue %d", epilogue);
  LSHM_REQUIRE(epilogue != LSHM_EPI_DELU || aux != nullptr, "lshm_up1d: DELU epilogue needs aux");
  LSHM_REQUIRE((reinterpret_cast<uintptr_t>(wimg) & 15) == 0, "lshm_up1d: weight image must be 16-byte aligned");
  LSHM_REQUIRE(pad == 1 || ((reinterpret_cast<uintptr_t>(big) & 15) == 0 && (big_ns & 3) == 0),
               "lshm_up1d: output must be 16-byte aligned for pad=0");
  if (N == 0) return LSHM_OK;
  UpArgs a{};
  a.small_ = small_; a.small_ns = small_ns; a.wimg = reinterpret_cast<const uint8_t*>(wimg); a.bias = bias;
  a.aux = epilogue == LSHM_EPI_DELU ? aux : nullptr; a.aux_ns = aux_ns; a.big = big; a.big_ns = big_ns;
  a.N = N; a.A = A; a.Bc = Bc; a.h = 1; a.w = l; a.pad = pad; a.epi = epilogue;
  return launch_up(1, a, as_stream(stream));
}

}  // extern "C"
