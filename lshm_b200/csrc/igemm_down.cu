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
// Warp roles: 4 epilogue warps, 4 (or 2 x 4) producer warps, 1 MMA-issuing warp, 1 weight-loading warp;
// a ring of shared-memory stages (mbarrier full/empty), two fp32 accumulator sets in TMEM, bias / ELU /
// ELU' fused in the epilogue which writes NCHW / NCL directly.
#include <stdlib.h>
#include "conv_geom.cuh"
#include "tma.cuh"

namespace lshm {
namespace {

using namespace tc;

struct DownArgs {
  // operand planes (tma.cuh): tensor maps over the hi / lo halves; used by the PRE instances only
  alignas(64) CUtensorMap tm_hi;
  alignas(64) CUtensorMap tm_lo;
  const float* big; int64_t big_ns;
  const uint8_t* wimg;
  const float* bias;
  const float* aux; int64_t aux_ns;
  float* small_; int64_t small_ns;
  int64_t N; int A; int Bc; int h; int w; int pad; int epi;
  int slots; int nstage; int64_t Q; int64_t mtiles; int ntn; int wres_on;
  FastDiv d_pp, d_pw, d_w, d_ntn;   // divisors (h+1)(w+1), w+1, w, ntn
};

// warps 0-3 epilogue, 4-7 producers (group 0), 8 MMA issuer, 9 weight loader, 10-13 producers (group 1, G = 2)
constexpr int down_threads(int g) { return 320 + 128 * (g - 1); }
constexpr int MAXST = 6;

// Persistent, warp-specialised: every CTA walks work items (position tile x channel tile); the
// producers run ahead through a ring of shared-memory stages, the MMA warp alternates between two
// TMEM accumulators, the epilogue warps drain one while the next is being computed.
//
// One accumulator tile -> bias / activation -> global for the thread's output position.  Independent
// loads first (bias, ELU' operand), then the accumulator, then math and stores, in halves of 8 channels
// behind warp-uniform branches: the 8- and 12-channel layers (the two largest maps) skip the padding
// half.  The first version went channel by channel with expm1f and 64-bit index math (~1500
// instructions per tile for a lone warp, the bottleneck of the whole kernel); the epilogue mode is a
// template parameter so no element carries the predicated-off instructions of the other modes.
template <int NT, int EPI>
__device__ __forceinline__ void down_epilogue_tile(const DownArgs& a, uint32_t trow, int nt, float* outp,
                                                   const float* auxp, int64_t hw, bool ok) {
  // eight channels at a time (24 live data registers: the kernel runs at a 64-register cap and the
  // 16-wide version spilled)
#pragma unroll 1
  for (int h8 = 0; h8 < NT / 8; ++h8) {
    const int ch0 = nt * NT + h8 * 8;
    const int nch = min(8, a.A - ch0);
    if (nch <= 0) break;                      // warp-uniform
    float bs[8], ax[8];
    const float* bp = a.bias + ch0;
    const float* xp = auxp + (int64_t)ch0 * hw;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      bs[j] = (a.bias != nullptr && j < nch) ? __ldg(bp + j) : 0.f;
      ax[j] = 0.f;
      if (EPI == LSHM_EPI_DELU && ok && j < nch) ax[j] = __ldg(xp + (int64_t)j * hw);
    }
    float v[8];
    tmem_ld8(trow + h8 * 8, v);
    float* op = outp + (int64_t)ch0 * hw;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float r = v[j] + bs[j];
      if (EPI == LSHM_EPI_ELU) r = elu_fast(r);
      else if (EPI == LSHM_EPI_DELU) r *= delu_from_out(ax[j]);
      if (ok && j < nch) op[(int64_t)j * hw] = r;
    }
  }
}

// G = producer groups of four warps.  G = 2 (layers with more than one K block: the deep, small maps):
// the groups fill alternate stages, so two gather -> convert -> hand-off chains are in flight per CTA; those
// layers are bound by the latency of that chain (2-3.5 us per K block, one or two tiles per CTA), not by
// bandwidth or issue slots.
// PRE: the input arrives as operand planes (already space-to-depth, already split into bf16 hi / lo): the producer
// warps are replaced by ONE thread that issues two tensor-TMA box loads per stage (cp.async.bulk.tensor.3d).
template <int DIM, int NT, int KC, int G, bool PRE>
__global__ void __launch_bounds__(down_threads(G), (G == 1 ? 3 : (G == 2 ? 2 : 1))) igemm_down_kernel(const __grid_constant__ DownArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t full_bar[MAXST], empty_bar[MAXST], acc_full[2], acc_empty[2], w_bar;
  __shared__ uint32_t tmem_base;
  constexpr int T = DIM == 2 ? 4 : 1;
  constexpr int CC = KC / 8;
  constexpr uint32_t IMG = 2u * T * CC * NT * 16;
  constexpr uint32_t TMEM_COLS = 2 * NT <= 32 ? 32 : (2 * NT <= 64 ? 64 : (2 * NT <= 128 ? 128 : 256));
  const int SLOTS = a.slots, NS = a.nstage;
  constexpr int NSLOT = DIM == 2 ? 2 : 1;          // staged slots per producer thread (1-D tiles: 128 slots)
  const uint32_t zbytes = (uint32_t)CC * SLOTS * 16;
  // (operand-plane instances with a resident weight image keep no image slot in the stages)
  const bool wres_early = ((4 * a.Bc + KC - 1) / KC) == 1 && a.ntn == 1 && a.wres_on;
  const uint32_t stage_bytes = 2 * zbytes + ((PRE && wres_early) ? 0u : IMG);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int Kc = 4 * a.Bc;
  const int KB = (Kc + KC - 1) / KC;
  const int PW = a.w + 1, PH = a.h + 1;
  const int64_t total = a.mtiles * a.ntn;
  // one K block and one channel tile: the weight image is the same for every work item, so it is
  // fetched once per CTA and stays resident behind the stage ring
  const bool wres = KB == 1 && a.ntn == 1 && a.wres_on;
  uint8_t* wres_ptr = smem + (size_t)NS * stage_bytes;

  if (warp == 8) tmem_alloc(&tmem_base, TMEM_COLS);
  if (tid == 0) {
    // arrivals per stage: the four producer warps (or the one TMA-issuing thread) + the weight loader unless resident
    const uint32_t nprod = PRE ? 1u : 4u;
    for (int s = 0; s < MAXST; ++s) { mbar_init(&full_bar[s], wres ? nprod : nprod + 1); mbar_init(&empty_bar[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], 4); }
    mbar_init(&w_bar, 1);
    mbar_init_fence();
  }
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem = tmem_base;

  if (PRE && ((warp >= 4 && warp < 8) || warp >= 10)) {
    // ------------------------------------------------ producer = the copy engine: one box per half and stage
    if (warp == 4 && lane == 0) {
      tma_prefetch_desc(&a.tm_hi);
      tma_prefetch_desc(&a.tm_lo);
      Ring ring{0, 0};
      for (int64_t item = blockIdx.x; item < total; item += gridDim.x) {
        const int q0 = (int)(fdiv((uint32_t)item, a.d_ntn) * 128u);
        for (int kb = 0; kb < KB; ++kb, ring.next(NS)) {
          const int s = ring.s;
          uint8_t* zhi = smem + (size_t)s * stage_bytes;
          mbar_wait(&empty_bar[s], ring.ph ^ 1);
          mbar_arrive_expect_tx(&full_bar[s], 2 * zbytes);
          tma_load_3d(zhi, &a.tm_hi, 0, q0 / PLANE_ROW, kb * CC, &full_bar[s]);
          tma_load_3d(zhi + zbytes, &a.tm_lo, 0, q0 / PLANE_ROW, kb * CC, &full_bar[s]);
        }
      }
    }
  } else if ((warp >= 4 && warp < 8) || warp >= 10) {
    // ------------------------------------------------ producers: stage the Z tile (hi/lo bf16)
    const int grp = warp >= 10 ? 1 + ((warp - 10) >> 2) : 0;   // producer group
    const int ptid = (tid & 127);                       // 0..127 within the group (warps 4-7 / 10-13... tid-128, tid-320)
    const int H = 2 * a.h, W = 2 * a.w;
    const int64_t HW = (int64_t)H * W;
    // Register-staged path.  Per slot the source window is described once (base pointer of the first
    // channel, validity of the two rows / two columns); a chunk column is 8 loads at fixed offsets.
    Ring ring{0, 0};
    uint32_t unit = 0;                                  // (item, K block) counter: group g fills units u % G == g
    for (int64_t item = blockIdx.x; item < total; item += gridDim.x) {
      const int64_t q0 = (int64_t)fdiv((uint32_t)item, a.d_ntn) * 128;
      const float* sp[NSLOT]; bool sv[NSLOT], full[NSLOT], r0ok[NSLOT], r1ok[NSLOT], c0ok[NSLOT], c1ok[NSLOT]; int sj[NSLOT];
#pragma unroll
      for (int i = 0; i < NSLOT; ++i) {
        const int s = ptid + i * 128;
        const int64_t q = q0 + s;
        sv[i] = s < SLOTS && q < a.Q;
        sp[i] = a.big; full[i] = false; r0ok[i] = r1ok[i] = c0ok[i] = c1ok[i] = false; sj[i] = 0;
        if (sv[i]) {
          const uint32_t uq = (uint32_t)q;            // Q < 2^31 (launcher): 32-bit divisions only
          if (DIM == 2) {
            const uint32_t n = fdiv(uq, a.d_pp);
            const uint32_t r = uq - n * (uint32_t)(PH * PW);
            const int by = (int)fdiv(r, a.d_pw), bx = (int)r - by * PW;
            r0ok[i] = by > 0; r1ok[i] = by < a.h; c0ok[i] = bx > 0; c1ok[i] = bx < a.w;
            full[i] = r0ok[i] && r1ok[i] && c0ok[i] && c1ok[i];
            sp[i] = a.big + (int64_t)n * a.big_ns + (int64_t)(2 * by - 1) * W + (2 * bx - 1);
          } else {
            const uint32_t n = fdiv(uq, a.d_w);       // 1-D: a.w holds the small length l
            sj[i] = (int)(uq - n * (uint32_t)a.w);
            sp[i] = a.big + (int64_t)n * a.big_ns + 4 * (int64_t)sj[i] - a.pad;
          }
        }
      }
      for (int kb = 0; kb < KB; ++kb, ring.next(NS), ++unit) {
        if (G > 1 && (int)(unit & (G - 1)) != grp) continue;   // another group's stage
        const int s = ring.s;
        uint8_t* zhi = smem + (size_t)s * stage_bytes;
        uint8_t* zlo = zhi + zbytes;
        const int ccb = (min(KC, Kc - kb * KC)) >> 3;
        mbar_wait(&empty_bar[s], ring.ph ^ 1);
        // per slot: fetch every chunk column, then convert (32 data registers -> 3 CTAs per SM)
#pragma unroll
        for (int i = 0; i < NSLOT; ++i) {
          const int slot = ptid + i * 128;
          if (slot >= SLOTS) continue;
          float v[CC][8];
#pragma unroll
          for (int cc = 0; cc < CC; ++cc) {
#pragma unroll
            for (int e = 0; e < 8; ++e) v[cc][e] = 0.f;
            const int b0 = 2 * (kb * CC + cc);          // Bc % 4 == 0: a chunk column is all-valid or all-padding
            if (cc < ccb && sv[i] && b0 < a.Bc) {
              if (DIM == 2) {
                const float* p0 = sp[i] + (int64_t)b0 * HW;
                const float* p1 = p0 + HW;
                if (full[i]) {
                  v[cc][0] = __ldg(p0); v[cc][1] = __ldg(p0 + 1); v[cc][2] = __ldg(p0 + W); v[cc][3] = __ldg(p0 + W + 1);
                  v[cc][4] = __ldg(p1); v[cc][5] = __ldg(p1 + 1); v[cc][6] = __ldg(p1 + W); v[cc][7] = __ldg(p1 + W + 1);
                } else {
                  if (r0ok[i] && c0ok[i]) { v[cc][0] = __ldg(p0); v[cc][4] = __ldg(p1); }
                  if (r0ok[i] && c1ok[i]) { v[cc][1] = __ldg(p0 + 1); v[cc][5] = __ldg(p1 + 1); }
                  if (r1ok[i] && c0ok[i]) { v[cc][2] = __ldg(p0 + W); v[cc][6] = __ldg(p1 + W); }
                  if (r1ok[i] && c1ok[i]) { v[cc][3] = __ldg(p0 + W + 1); v[cc][7] = __ldg(p1 + W + 1); }
                }
              } else {
                const int64_t Lb = 4 * (int64_t)a.w;
                const float* p0 = sp[i] + (int64_t)b0 * Lb;
                const float* p1 = p0 + Lb;
                if (a.pad == 0) {
                  const float4 x0 = __ldg(reinterpret_cast<const float4*>(p0));
                  const float4 x1 = __ldg(reinterpret_cast<const float4*>(p1));
                  v[cc][0] = x0.x; v[cc][1] = x0.y; v[cc][2] = x0.z; v[cc][3] = x0.w;
                  v[cc][4] = x1.x; v[cc][5] = x1.y; v[cc][6] = x1.z; v[cc][7] = x1.w;
                } else {
                  if (sj[i] > 0) { v[cc][0] = __ldg(p0); v[cc][4] = __ldg(p1); }
                  v[cc][1] = __ldg(p0 + 1); v[cc][2] = __ldg(p0 + 2); v[cc][3] = __ldg(p0 + 3);
                  v[cc][5] = __ldg(p1 + 1); v[cc][6] = __ldg(p1 + 2); v[cc][7] = __ldg(p1 + 3);
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
    }
  } else if (warp < 4) {
    // ------------------------------------------------ epilogue: TMEM -> bias/act -> global
    const int64_t hw = DIM == 2 ? (int64_t)a.h * a.w : (int64_t)a.w;
    uint32_t tc_ = 0;
    for (int64_t item = blockIdx.x; item < total; item += gridDim.x, ++tc_) {
      const uint32_t mt = fdiv((uint32_t)item, a.d_ntn);
      const int nt = (int)((uint32_t)item - mt * (uint32_t)a.ntn);
      const int64_t q = (int64_t)mt * 128 + tid;
      bool ok = q < a.Q;
      int64_t n = 0, pos = 0;
      if (ok) {
        const uint32_t uq = (uint32_t)q;
        if (DIM == 2) {
          const uint32_t un = fdiv(uq, a.d_pp);
          const uint32_t r = uq - un * (uint32_t)(PH * PW);
          const int by = (int)fdiv(r, a.d_pw), bx = (int)r - by * PW;
          ok = by < a.h && bx < a.w;
          n = un; pos = (int64_t)by * a.w + bx;
        } else {
          const uint32_t un = fdiv(uq, a.d_w);
          n = un; pos = uq - un * (uint32_t)a.w;
        }
      }
      float* outp = a.small_ + n * a.small_ns + pos;
      const float* auxp = a.aux != nullptr ? a.aux + n * a.aux_ns + pos : nullptr;
      const uint32_t buf = tc_ & 1;
      mbar_wait(&acc_full[buf], (tc_ >> 1) & 1);
      fence_after();
      const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16) + buf * NT;
      if (a.epi == LSHM_EPI_ELU) down_epilogue_tile<NT, LSHM_EPI_ELU>(a, trow, nt, outp, auxp, hw, ok);
      else if (a.epi == LSHM_EPI_DELU) down_epilogue_tile<NT, LSHM_EPI_DELU>(a, trow, nt, outp, auxp, hw, ok);
      else down_epilogue_tile<NT, LSHM_EPI_NONE>(a, trow, nt, outp, auxp, hw, ok);
      fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[buf]);
    }
  } else if (warp == 8) {
    // ------------------------------------------------ MMA issuer: warp-uniform loop, the elected lane issues
    // (under `if (lane == 0)` every tcgen05.mma was wrapped in a vote/broadcast waterfall, ~16
    // instructions per MMA; in uniform code the descriptors stay in uniform registers)
    const uint32_t idesc = make_idesc(NT, 0, 0);
    const uint32_t leader = elect_one();
    uint32_t tc_ = 0;
    Ring ring{0, 0};
    if (wres) mbar_wait(&w_bar, 0);
    for (int64_t item = blockIdx.x; item < total; item += gridDim.x, ++tc_) {
      const uint32_t buf = tc_ & 1;
      mbar_wait(&acc_empty[buf], ((tc_ >> 1) & 1) ^ 1);
      fence_after();
      uint32_t acc = 0;
      for (int kb = 0; kb < KB; ++kb, ring.next(NS)) {
        const int s = ring.s;
        mbar_wait(&full_bar[s], ring.ph);
        fence_after();
        const uint32_t zhi = smem_u32(smem + (size_t)s * stage_bytes);
        const uint32_t bhi = wres ? smem_u32(wres_ptr) : zhi + 2 * zbytes;
        const uint64_t dah = make_desc(zhi, SLOTS * 16, 128), dal = make_desc(zhi + zbytes, SLOTS * 16, 128);
        const uint64_t dbh = make_desc(bhi, NT * 16, 128), dbl = make_desc(bhi + IMG / 2, NT * 16, 128);
        const int ksteps = (min(KC, Kc - kb * KC)) >> 4;
        // rolled on purpose: the issuing warp's code must stay small (instruction-cache footprint)
#pragma unroll 1
        for (int tap = 0; tap < T; ++tap) {
          const uint32_t shift = DIM == 2 ? (uint32_t)((tap >> 1) * PW + (tap & 1)) : 0u;
#pragma unroll 1
          for (int ks = 0; ks < ksteps; ++ks) {
            const uint32_t ao = (uint32_t)(2 * ks) * SLOTS + shift;            // 16-byte units
            const uint32_t bo = (uint32_t)(tap * CC + 2 * ks) * NT;
            mma_split3_warp(tmem + buf * NT, desc_off(dah, ao), desc_off(dal, ao), desc_off(dbh, bo), desc_off(dbl, bo),
                            idesc, acc, leader);
            acc = 1;
          }
        }
        commit_warp(&empty_bar[s], leader);
      }
      commit_warp(&acc_full[buf], leader);
    }
  } else {
    // ------------------------------------------------ loader warp: weight images
    if (wres) {
      // resident image: one fetch, the stage barriers then count the four producer warps only
      if (lane == 0) {
        mbar_arrive_expect_tx(&w_bar, IMG);
        bulk_g2s(wres_ptr, a.wimg, IMG, &w_bar);
      }
    } else {
      Ring ring{0, 0};
      for (int64_t item = blockIdx.x; item < total; item += gridDim.x) {
        const uint32_t mt = fdiv((uint32_t)item, a.d_ntn);
        const int nt = (int)((uint32_t)item - mt * (uint32_t)a.ntn);
        for (int kb = 0; kb < KB; ++kb, ring.next(NS)) {
          const int s = ring.s;
          mbar_wait(&empty_bar[s], ring.ph ^ 1);
          uint8_t* stage = smem + (size_t)s * stage_bytes;
          if (lane == 0) {
            mbar_arrive_expect_tx(&full_bar[s], IMG);
            bulk_g2s(stage + 2 * zbytes, a.wimg + ((size_t)nt * KB + kb) * IMG, IMG, &full_bar[s]);
          }
        }
      }
    }
  }
  fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tmem, TMEM_COLS);
}

template <int DIM, int NT, int KC, int G, bool PRE = false>
int launch_down_t(DownArgs a, const DownGeom& g, cudaStream_t st) {
  const bool wres = g.KB == 1 && g.ntiles == 1 && a.wres_on;
  const size_t stage = (size_t)2 * (KC / 8) * a.slots * 16 + ((PRE && wres) ? 0 : g.img);
  const int64_t units = a.mtiles * g.ntiles * g.KB;
  // two CTAs per SM when the ring fits in ~110 KB each (227 KB per SM), else one CTA with a deep ring
  // three CTAs per SM when two stages fit in ~74 KB, else two, else one with a deep ring
  const size_t third_sm = 74 * 1024, half_sm = 110 * 1024, full_sm = 200 * 1024;
  int ns = (int)((third_sm - g.img) / stage);
  if (ns < 2) ns = (int)((half_sm - g.img) / stage);
  if (ns < 2) ns = (int)((full_sm - g.img) / stage);
  // Few work items, many K blocks (the deep layers: one or two items per SM, 12-24 K blocks each): several CTAs per SM
  // buy nothing, but with two stages every K block waits for its weight image (a 12-25 KB bulk copy from L2, ~1 us of
  // latency): one CTA per SM with the deepest ring instead.
  static const bool no_deep_ring = getenv("LSHM_DOWN_NODEEP") != nullptr;   // experiment switch
  if (!no_deep_ring && g.KB >= 4 && a.mtiles * g.ntiles <= 2 * (int64_t)sm_count())
    ns = std::max(ns, (int)((full_sm - g.img) / stage));
  ns = std::min(ns, MAXST);
  LSHM_REQUIRE(ns >= 1, "igemm_down: tile does not fit in shared memory");
  a.nstage = (int)std::min<int64_t>(ns, std::max<int64_t>(1, units));
  const size_t smem = stage * a.nstage + g.img;   // + resident weight image (used when KB == ntiles == 1)
  if (PRE) {
    const PlaneGeom pg = plane_geom(DIM, a.N, a.Bc, a.h, a.w);
    const uint8_t* base = reinterpret_cast<const uint8_t*>(a.big);
    if (int rc = make_plane_tmap(&a.tm_hi, base, pg.Qs, pg.chunks, a.slots, KC / 8)) return rc;
    if (int rc = make_plane_tmap(&a.tm_lo, base + pg.half_bytes, pg.Qs, pg.chunks, a.slots, KC / 8)) return rc;
  }
  LSHM_CUDA(cudaFuncSetAttribute(igemm_down_kernel<DIM, NT, KC, G, PRE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "igemm_down");
  const int per_sm = std::min(G == 1 ? 3 : (G == 2 ? 2 : 1), smem <= 74 * 1024 ? 3 : (smem <= 112 * 1024 ? 2 : 1));
  const int64_t grid = std::min<int64_t>(a.mtiles * g.ntiles, (int64_t)sm_count() * per_sm);
  igemm_down_kernel<DIM, NT, KC, G, PRE><<<(unsigned)grid, down_threads(G), smem, st>>>(a);
  LSHM_CHECK_LAUNCH("igemm_down");
  return LSHM_OK;
}

int launch_down(int dim, DownArgs a, cudaStream_t st, bool planes = false) {
  const DownGeom g = down_geom(dim, a.A, a.Bc);
  a.slots = dim == 2 ? (128 + a.w + 2 + 7) / 8 * 8 : 128;
  if (planes) a.slots = (a.slots + PLANE_ROW - 1) / PLANE_ROW * PLANE_ROW;   // whole 512-byte rows of the tensor map
  a.Q = dim == 2 ? a.N * (int64_t)(a.h + 1) * (a.w + 1) : a.N * (int64_t)a.w;
  a.d_pp = make_fastdiv((uint32_t)((a.h + 1) * (a.w + 1))); a.d_pw = make_fastdiv((uint32_t)(a.w + 1)); a.d_w = make_fastdiv((uint32_t)a.w);
  LSHM_REQUIRE(a.Q < (1LL << 31) - 4096, "lshm_down: too many positions (%lld) for one call; split the batch", (long long)a.Q);
  a.mtiles = ceil_div(a.Q, 128);
  a.ntn = g.ntiles;
  a.d_ntn = make_fastdiv((uint32_t)a.ntn);
  static const bool wres_off = getenv("LSHM_DOWN_NOWRES") != nullptr;   // experiment switch
  a.wres_on = wres_off ? 0 : 1;
  static const bool one_group = getenv("LSHM_DOWN_G1") != nullptr;       // experiment switch
  const bool two = g.KB >= 2 && !one_group;                              // two producer groups (see the kernel)
  if (planes) {
    // the copy engine is the producer: any layer whose input exists as operand planes
#define LP(D, NTV, KCV) return launch_down_t<D, NTV, KCV, 1, true>(a, g, st)
    if (dim == 2) {
      switch (g.NT) {
        case 16: LP(2, 16, 32);
        case 32: LP(2, 32, 32);
        case 48: LP(2, 48, 32);
        default: LP(2, 96, 16);
      }
    } else {
      switch (g.NT) {
        case 16: LP(1, 16, 32);
        case 32: LP(1, 32, 32);
        case 48: LP(1, 48, 32);
        default: LP(1, 96, 32);
      }
    }
#undef LP
  }
  // four producer groups for the deepest layers (one or two items per SM, >= 8 K blocks each: the kernel is a chain of
  // gather -> convert -> hand-off latencies, four of them in flight instead of two)
  static const bool no_g4 = getenv("LSHM_DOWN_NOG4") != nullptr;          // experiment switch
  const bool four = two && !no_g4 && g.NT == 96 && g.KB >= 8 && a.mtiles * g.ntiles <= 2 * (int64_t)sm_count();
#define LD(D, NTV, KCV) do { if (two) return launch_down_t<D, NTV, KCV, 2>(a, g, st); return launch_down_t<D, NTV, KCV, 1>(a, g, st); } while (0)
  if (four) {
    if (dim == 2) return launch_down_t<2, 96, 16, 4>(a, g, st);
    return launch_down_t<1, 96, 32, 4>(a, g, st);
  }
  if (dim == 2) {
    switch (g.NT) {
      case 16: LD(2, 16, 32);
      case 32: LD(2, 32, 32);
      case 48: LD(2, 48, 32);
      default: LD(2, 96, 16);
    }
  } else {
    switch (g.NT) {
      case 16: LD(1, 16, 32);
      case 32: LD(1, 32, 32);
      case 48: LD(1, 48, 32);
      default: LD(1, 96, 32);
    }
  }
#undef LD
}

}  // namespace

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
  // the producers stage 2 x 128 slots per tile (tile + one halo row): 128 + w + 2 <= 256
  LSHM_REQUIRE(w_ <= 126, "lshm_down2d: small-map width %d too large (max 126)", w_);
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

// Same kernels with the input given as operand planes (lshm_stage_planes*, lshm_cascade_combine_planes,
// lshm_residual_split_planes): the tiles are fetched by tensor-TMA box loads instead of producer warps.
int lshm_down2d_planes(const void* planes, const void* wimg, const float* bias,
                       const float* aux, int64_t aux_ns, float* small_, int64_t small_ns,
                       int64_t N, int A, int Bc, int h, int w_, int epilogue, lshm_stream_t stream) {
  LSHM_REQUIRE(planes && wimg && small_, "lshm_down2d_planes: null pointer");
  LSHM_REQUIRE(N >= 0 && A > 0 && Bc > 0 && (Bc & 3) == 0 && h > 0 && w_ > 0, "lshm_down2d_planes: bad sizes (Bc must be a multiple of 4)");
  LSHM_REQUIRE(epilogue >= 0 && epilogue <= 2, "lshm_down2d_planes: bad epilogue %d", epilogue);
  LSHM_REQUIRE(epilogue != LSHM_EPI_DELU || aux != nullptr, "lshm_down2d_planes: DELU epilogue needs aux");
  LSHM_REQUIRE(w_ <= 126, "lshm_down2d_planes: small-map width %d too large (max 126)", w_);
  LSHM_REQUIRE((reinterpret_cast<uintptr_t>(wimg) & 15) == 0, "lshm_down2d_planes: weight image must be 16-byte aligned");
  if (N == 0) return LSHM_OK;
  DownArgs a{};
  a.big = reinterpret_cast<const float*>(planes); a.big_ns = 0; a.wimg = reinterpret_cast<const uint8_t*>(wimg); a.bias = bias;
  a.aux = epilogue == LSHM_EPI_DELU ? aux : nullptr; a.aux_ns = aux_ns;
  a.small_ = small_; a.small_ns = small_ns; a.N = N; a.A = A; a.Bc = Bc; a.h = h; a.w = w_; a.pad = 0; a.epi = epilogue;
  return launch_down(2, a, as_stream(stream), true);
}

int lshm_down1d_planes(const void* planes, const void* wimg, const float* bias,
                       const float* aux, int64_t aux_ns, float* small_, int64_t small_ns,
                       int64_t N, int A, int Bc, int l, int epilogue, lshm_stream_t stream) {
  LSHM_REQUIRE(planes && wimg && small_, "lshm_down1d_planes: null pointer");
  LSHM_REQUIRE(N >= 0 && A > 0 && Bc > 0 && (Bc & 3) == 0 && l > 0, "lshm_down1d_planes: bad sizes (Bc must be a multiple of 4)");
  LSHM_REQUIRE(epilogue >= 0 && epilogue <= 2, "lshm_down1d_planes: bad epilogue %d", epilogue);
  LSHM_REQUIRE(epilogue != LSHM_EPI_DELU || aux != nullptr, "lshm_down1d_planes: DELU epilogue needs aux");
  LSHM_REQUIRE((reinterpret_cast<uintptr_t>(wimg) & 15) == 0, "lshm_down1d_planes: weight image must be 16-byte aligned");
  if (N == 0) return LSHM_OK;
  DownArgs a{};
  a.big = reinterpret_cast<const float*>(planes); a.big_ns = 0; a.wimg = reinterpret_cast<const uint8_t*>(wimg); a.bias = bias;
  a.aux = epilogue == LSHM_EPI_DELU ? aux : nullptr; a.aux_ns = aux_ns;
  a.small_ = small_; a.small_ns = small_ns; a.N = N; a.A = A; a.Bc = Bc; a.h = 1; a.w = l; a.pad = 0; a.epi = epilogue;
  return launch_down(1, a, as_stream(stream), true);
}

}  // extern "C"
