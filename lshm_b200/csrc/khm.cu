// K-harmonic means on latent vectors: loss, analytic gradient, assignment, centre sums.
//
// Reference semantics: Kmeans.forward  /root/reference/src/lofar_models.py:199-209,
// offline_update :231-261 (intent), eval loop src/evaluate_clustering.py:110-119.
//
// Design (B200): the direct-difference form sum_l (x_l - m_kl)^2 is kept in fp32 (the GEMM
// form cancels exactly where the harmonic mean is dominated).  A point is held in
// registers by TPP adjacent lanes (<= 32 floats per lane, interleaved float4 chunks so the lanes
// of one point read consecutive 16-byte words of the row), centres live in shared memory and are
// read as broadcast float4; the per-point harmonic sum is reduced across the TPP lanes with
// warp shuffles.  The centre-gradient / centre-update sums are a [K x pts] x [pts x L] product,
// done per tile from shared memory with one owner thread per output (no atomics inside the tile
// loop); blocks are persistent (grid = multiple of the SM count) and flush once.
#include <stdlib.h>
#include "common.cuh"

namespace lshm {
namespace {

constexpr int KHM_THREADS = 128;
constexpr int KHM_KC = 32;        // centres per staged chunk (streaming mode) / per w-tile
constexpr int KHM_MAXCH = 8;      // float4 chunks per lane  (<= 32 floats of a point per lane)
constexpr float KHM_EPS = 1e-9f;  // src/lofar_models.py:195

struct KhmArgs {
  const float* X; int64_t ldx; const float* M; int64_t N; int K; int L;
  float p; int pmode;                       // pmode: 2, 4 or 0 (generic powf)
  // pass 1 outputs
  double* loss_sum; float* e_out; int32_t* ids; float* dist; int group;
  // pass 2
  float gscale; float* gX; int64_t ldg; int accumulate_x; float* gM;  // BWD
  float* num; float* den;                                              // SUMS
};

__device__ __forceinline__ float pow_p(float d2, float p, int pmode) {    // d^p from d^2
  if (pmode == 2) return d2;
  if (pmode == 4) return d2 * d2;
  return powf(d2, 0.5f * p);
}
__device__ __forceinline__ float pow_pm2(float d2, float p, int pmode) {  // d^(p-2)
  if (pmode == 2) return 1.f;
  if (pmode == 4) return d2;
  return powf(d2, 0.5f * (p - 2.f));
}

// ---- packed fp32x2 arithmetic (sm_100 FFMA2 / FADD2): halves the issue slots of the distance loop,
//      which is what bounds the K <= 16 (HBM-side) cases
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float a, float b) { f32x2 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk2(f32x2 v, float& a, float& b) { asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b) { f32x2 r; asm("sub.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) { f32x2 r; asm("mul.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }

// acc += a * b in place (the "+l" form keeps the accumulator in its register pair: with the three-operand form the
// compiler wrote half of the gradient sums into the operand's registers and moved them back - 16 MOVs per centre)
__device__ __forceinline__ void fma2_acc(f32x2& acc, f32x2 a, f32x2 b) { asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(a), "l"(b)); }

// 1/x to 1 ulp in one MUFU op (the IEEE-rounded division is ~12 instructions per (point, centre) pair,
// a sixth of the K = 10 inner loop)
__device__ __forceinline__ float rcp_fast(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }

template <int TPP>
__device__ __forceinline__ float lanes_sum(float v) {
#pragma unroll
  for (int o = 1; o < TPP; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// squared distance between the lane's slice of x and the matching slice of centre row `mrow`
template <int TPP, int NCH>
__device__ __forceinline__ float dist2(const float4 (&x4)[KHM_MAXCH], const float* mrow, int s, int nch) {
  f32x2 a0 = 0ull, a1 = 0ull;   // two packed accumulators = four independent chains
#pragma unroll
  for (int c = 0; c < KHM_MAXCH; ++c) {
    if (NCH > 0 ? c < NCH : c < nch) {
      const float4 m = *reinterpret_cast<const float4*>(mrow + ((c * TPP + s) << 2));
      const f32x2 d0 = sub2(pk2(x4[c].x, x4[c].y), pk2(m.x, m.y));
      const f32x2 d1 = sub2(pk2(x4[c].z, x4[c].w), pk2(m.z, m.w));
      a0 = fma2(d0, d0, a0);
      a1 = fma2(d1, d1, a1);
    }
  }
  float p, q, r, t;
  upk2(a0, p, q); upk2(a1, r, t);
  return lanes_sum<TPP>((p + q) + (r + t));
}

// two points against one centre row: every broadcast LDS.128 of the centre feeds both points.  A warp-wide
// 16-byte shared load costs 4 crossbar cycles even when it is a broadcast, so with one point per thread
// the distance loop is shared-memory-bound at ~32 (point,k,l) elements per clock per SM (measured: 25).
template <int TPP, int NCH>
__device__ __forceinline__ void dist2x2(const float4 (&xa)[KHM_MAXCH], const float4 (&xb)[KHM_MAXCH], const float* mrow,
                                        int s, int nch, float& da, float& db) {
  f32x2 a0 = 0ull, a1 = 0ull, b0 = 0ull, b1 = 0ull;
#pragma unroll
  for (int c = 0; c < KHM_MAXCH; ++c) {
    if (NCH > 0 ? c < NCH : c < nch) {
      const float4 m = *reinterpret_cast<const float4*>(mrow + ((c * TPP + s) << 2));
      const f32x2 m0 = pk2(m.x, m.y), m1 = pk2(m.z, m.w);
      const f32x2 p0 = sub2(pk2(xa[c].x, xa[c].y), m0), p1 = sub2(pk2(xa[c].z, xa[c].w), m1);
      const f32x2 q0 = sub2(pk2(xb[c].x, xb[c].y), m0), q1 = sub2(pk2(xb[c].z, xb[c].w), m1);
      a0 = fma2(p0, p0, a0); a1 = fma2(p1, p1, a1);
      b0 = fma2(q0, q0, b0); b1 = fma2(q1, q1, b1);
    }
  }
  float p, q, r, t;
  upk2(a0, p, q); upk2(a1, r, t);
  da = lanes_sum<TPP>((p + q) + (r + t));
  upk2(b0, p, q); upk2(b1, r, t);
  db = lanes_sum<TPP>((p + q) + (r + t));
}

template <int TPP, int NCH>
__device__ __forceinline__ void load_point(float4 (&x4)[KHM_MAXCH], const float* X, int64_t ldx,
                                           int64_t i, bool valid, int s, int nch) {
#pragma unroll
  for (int c = 0; c < KHM_MAXCH; ++c) {
    x4[c] = make_float4(0.f, 0.f, 0.f, 0.f);
    if ((NCH > 0 ? c < NCH : c < nch) && valid) x4[c] = *reinterpret_cast<const float4*>(X + i * ldx + ((c * TPP + s) << 2));
  }
}

__device__ __forceinline__ void stage_centres(float* ms, const float* M, int k0, int kc, int L) {
  const int n4 = (kc * L) >> 2;
  const float4* src = reinterpret_cast<const float4*>(M + (int64_t)k0 * L);
  float4* dst = reinterpret_cast<float4*>(ms);
  for (int i = threadIdx.x; i < n4; i += blockDim.x) dst[i] = src[i];
}

// ------------------------------------------------------------------------------------------
// pass 1: e_i = sum_k 1/(d_ik^p + eps); loss, e_out, argmin ids, group distances
// ------------------------------------------------------------------------------------------
// NCH = KHM_MAXCH: every lane holds all eight float4 chunks (L = 32*TPP), the chunk loops carry no
// guards (the guarded form cost a branch per chunk plus 8 accumulator moves: 36% useful FFMA2/FADD2 in
// the K=10, L=64 capture); NCH = 0: chunk count known at run time only.
// MODE 0: harmonic sums only (forward loss); 1: nearest centre only (assignment); 2: everything (group
// distances).  The per-centre tail is ~a third of the K = 10 inner loop, so the two hot entry points do not
// carry each other's part.
template <int TPP, bool RESIDENT, int NCH, int MODE>
__global__ void __launch_bounds__(KHM_THREADS) khm_pass1_kernel(KhmArgs a) {
  extern __shared__ __align__(16) float smem[];
  float* ms = smem;  // RESIDENT: K*L, else KC*L
  __shared__ double red[32];
  constexpr int PTS = KHM_THREADS / TPP;        // a tile is 2*PTS points: two per thread group
  const int L = a.L, K = a.K, nch = L / (4 * TPP);
  const int pt = threadIdx.x / TPP, s = threadIdx.x % TPP;
  const int64_t ntiles = (a.N + 2 * PTS - 1) / (2 * PTS);
  const float Kf = (float)K;
  double lsum = 0.0;
  if (RESIDENT) { stage_centres(ms, a.M, 0, K, L); __syncthreads(); }
  const float inv_group = a.group > 0 ? 1.f / (float)a.group : 0.f;
  for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const int64_t ia = t * 2 * PTS + pt, ib = ia + PTS;
    const bool va = ia < a.N, vb = ib < a.N;
    float4 xa[KHM_MAXCH], xb[KHM_MAXCH];
    if constexpr (TPP == 1 && NCH == KHM_MAXCH) {
      // L = 32, a whole point per lane: lane i reading row i directly makes every LDG.128 touch 32 different 128-byte
      // lines (8 load wavefronts per point, the bound of this case).  The warp's 32 rows are 4 KB contiguous when the
      // rows are dense: fetch them as consecutive 16-byte words (4 lines per instruction) and transpose through a
      // padded per-warp buffer (144-byte row pitch: both sides conflict-free).
      if (a.ldx == 32) {
        float* stg = smem + (RESIDENT ? K * L : KHM_KC * L) + (threadIdx.x >> 5) * (32 * 36);
        const int lane = threadIdx.x & 31;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int64_t row0 = t * 2 * PTS + h * PTS + (threadIdx.x & ~31);   // first row of the warp's 32
          float4 (&x4)[KHM_MAXCH] = h == 0 ? xa : xb;
          float4 w[KHM_MAXCH];
#pragma unroll
          for (int c = 0; c < KHM_MAXCH; ++c) {
            const int word = c * 32 + lane;                                  // 16-byte word of the 4 KB block
            w[c] = row0 + (word >> 3) < a.N ? *reinterpret_cast<const float4*>(a.X + row0 * 32 + word * 4)
                                            : make_float4(0.f, 0.f, 0.f, 0.f);
          }
          __syncwarp();
#pragma unroll
          for (int c = 0; c < KHM_MAXCH; ++c) {
            const int word = c * 32 + lane;
            *reinterpret_cast<float4*>(stg + (word >> 3) * 36 + (word & 7) * 4) = w[c];
          }
          __syncwarp();
#pragma unroll
          for (int c = 0; c < KHM_MAXCH; ++c) x4[c] = *reinterpret_cast<const float4*>(stg + lane * 36 + c * 4);
        }
      } else {
        load_point<TPP, NCH>(xa, a.X, a.ldx, ia, va, s, nch);
        load_point<TPP, NCH>(xb, a.X, a.ldx, ib, vb, s, nch);
      }
    } else {
      load_point<TPP, NCH>(xa, a.X, a.ldx, ia, va, s, nch);
      load_point<TPP, NCH>(xb, a.X, a.ldx, ib, vb, s, nch);
    }
    float ea = 0.f, eb = 0.f, besta = 3.4e38f, bestb = 3.4e38f;
    int bia = 0, bib = 0;
    for (int k0 = 0; k0 < K; k0 += (RESIDENT ? K : KHM_KC)) {
      const int kc = RESIDENT ? K : min(KHM_KC, K - k0);
      if (!RESIDENT) { __syncthreads(); stage_centres(ms, a.M, k0, kc, L); __syncthreads(); }
#pragma unroll 2
      for (int kk = 0; kk < kc; ++kk) {
        float da, db;
        dist2x2<TPP, NCH>(xa, xb, ms + kk * L, s, nch, da, db);
        float pa = 0.f, pb = 0.f;
        if (MODE != 1) {
          pa = pow_p(da, a.p, a.pmode); pb = pow_p(db, a.p, a.pmode);
          ea += rcp_fast(pa + KHM_EPS);
          eb += rcp_fast(pb + KHM_EPS);
        }
        if (MODE != 0) {
          if (da < besta) { besta = da; bia = k0 + kk; }
          if (db < bestb) { bestb = db; bib = k0 + kk; }
        }
        if (MODE == 2 && a.dist != nullptr && s == 0) {
          if (va) atomicAdd(a.dist + (ia / a.group) * K + (k0 + kk), pa * inv_group);
          if (vb) atomicAdd(a.dist + (ib / a.group) * K + (k0 + kk), pb * inv_group);
        }
      }
    }
    if (s == 0) {
      if (va) { lsum += (double)(Kf / (ea + KHM_EPS)); if (a.e_out) a.e_out[ia] = ea; if (a.ids) a.ids[ia] = bia; }
      if (vb) { lsum += (double)(Kf / (eb + KHM_EPS)); if (a.e_out) a.e_out[ib] = eb; if (a.ids) a.ids[ib] = bib; }
    }
  }
  if (a.loss_sum != nullptr) {
    const double tot = block_sum<double>(lsum, red);
    if (threadIdx.x == 0) atomicAdd(a.loss_sum, tot);
  }
}

// ------------------------------------------------------------------------------------------
// pass 2: weights w_ik (gradient) or Q_ik (centre update); gX in registers; [K,L] sums via a
// shared-memory tile product
// ------------------------------------------------------------------------------------------
template <int TPP, bool RESIDENT, bool SUMS, int NCH>
__global__ void __launch_bounds__(KHM_THREADS) khm_pass2_kernel(KhmArgs a) {
  extern __shared__ __align__(16) float smem[];
  constexpr int PTS = KHM_THREADS / TPP;
  const int L = a.L, K = a.K, nch = L / (4 * TPP), L4 = L >> 2;
  // smem carve-up
  float* ms = smem;                                       // RESIDENT: K*L else KC*L
  float* acc_s = ms + (RESIDENT ? K * L : KHM_KC * L);    // RESIDENT: K*L accumulators
  float* wsum_s = acc_s + (RESIDENT ? K * L : 0);         // RESIDENT: K
  float* xs = wsum_s + (RESIDENT ? ((K + 3) & ~3) : 0);   // PTS*L
  float* ws = xs + PTS * L;                               // PTS*KC
  // K <= KHM_KC: the squared distances of the harmonic-sum loop are kept (one float per point and centre)
  // for the weight loop instead of being recomputed - a quarter of the pass's arithmetic
  const bool cache_d2 = K <= KHM_KC;
  float* d2s = ws + PTS * KHM_KC;                         // [KC][PTS] when cache_d2
  __shared__ double red[32];
  const int pt = threadIdx.x / TPP, s = threadIdx.x % TPP;
  const int64_t ntiles = (a.N + PTS - 1) / PTS;
  const float Kf = (float)K;
  double lsum = 0.0;
  if (RESIDENT) {
    stage_centres(ms, a.M, 0, K, L);
    for (int i = threadIdx.x; i < K * L; i += blockDim.x) acc_s[i] = 0.f;
    for (int i = threadIdx.x; i < K; i += blockDim.x) wsum_s[i] = 0.f;
    __syncthreads();
  }
  for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const int64_t i = t * PTS + pt;
    const bool valid = i < a.N;
    float4 x4[KHM_MAXCH];
    load_point<TPP, NCH>(x4, a.X, a.ldx, i, valid, s, nch);
    // ---- harmonic sum e_i
    float e = 0.f;
    for (int k0 = 0; k0 < K; k0 += (RESIDENT ? K : KHM_KC)) {
      const int kc = RESIDENT ? K : min(KHM_KC, K - k0);
      if (!RESIDENT) { __syncthreads(); stage_centres(ms, a.M, k0, kc, L); __syncthreads(); }
      for (int kk = 0; kk < kc; ++kk) {
        const float d2 = dist2<TPP, NCH>(x4, ms + kk * L, s, nch);
        if (cache_d2) d2s[kk * PTS + pt] = d2;            // every lane of the point writes (and later reads) the same value
        e += rcp_fast(pow_p(d2, a.p, a.pmode) + KHM_EPS);
      }
    }
    if (valid && s == 0) lsum += (double)(Kf / (e + KHM_EPS));
    // per-point coefficient: gradient K/(e+eps)^2 ; centre update alpha = 1/(e^2+eps)
    const float coef = SUMS ? 1.0f / (e * e + KHM_EPS) : Kf / ((e + KHM_EPS) * (e + KHM_EPS));
    // ---- x tile to shared memory for the [K x pts] x [pts x L] product
#pragma unroll
    for (int c = 0; c < KHM_MAXCH; ++c)
      if (NCH > 0 ? c < NCH : c < nch) *reinterpret_cast<float4*>(xs + pt * L + ((c * TPP + s) << 2)) = x4[c];
    float4 g4[KHM_MAXCH];
#pragma unroll
    for (int c = 0; c < KHM_MAXCH; ++c) g4[c] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int k0 = 0; k0 < K; k0 += KHM_KC) {
      const int kc = min(KHM_KC, K - k0);
      if (!RESIDENT) { __syncthreads(); stage_centres(ms, a.M, k0, kc, L); __syncthreads(); }
      const float* mbase = RESIDENT ? ms + k0 * L : ms;
      for (int kk = 0; kk < kc; ++kk) {
        const float* mrow = mbase + kk * L;
        const float d2 = cache_d2 ? d2s[kk * PTS + pt] : dist2<TPP, NCH>(x4, mrow, s, nch);
        const float dp = pow_p(d2, a.p, a.pmode);
        float w;
        if (SUMS) {
          w = coef * rcp_fast(dp * d2 + KHM_EPS);                 // alpha_i / (d^(p+2) + eps)
        } else {
          const float tt = dp + KHM_EPS;
          w = d2 > 0.f ? coef * a.p * pow_pm2(d2, a.p, a.pmode) * rcp_fast(tt * tt) : 0.f;
        }
        if (!valid) w = 0.f;
        if (!SUMS) {
#pragma unroll
          for (int c = 0; c < KHM_MAXCH; ++c) {
            if (NCH > 0 ? c < NCH : c < nch) {
              const float4 m = *reinterpret_cast<const float4*>(mrow + ((c * TPP + s) << 2));
              const f32x2 w2 = pk2(w, w);
              const f32x2 r0 = fma2(w2, sub2(pk2(x4[c].x, x4[c].y), pk2(m.x, m.y)), pk2(g4[c].x, g4[c].y));
              const f32x2 r1 = fma2(w2, sub2(pk2(x4[c].z, x4[c].w), pk2(m.z, m.w)), pk2(g4[c].z, g4[c].w));
              upk2(r0, g4[c].x, g4[c].y);
              upk2(r1, g4[c].z, g4[c].w);
            }
          }
        }
        if (s == 0) ws[pt * KHM_KC + kk] = w;
      }
      __syncthreads();
      // tile product: thread owns output (kk, l4)
      for (int o = threadIdx.x; o < kc * L4; o += blockDim.x) {
        const int kk = o / L4, l4 = o - kk * L4;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        float wsum = 0.f;
#pragma unroll 4
        for (int q = 0; q < PTS; ++q) {
          const float wv = ws[q * KHM_KC + kk];
          const float4 xv = *reinterpret_cast<const float4*>(xs + q * L + (l4 << 2));
          acc.x = fmaf(wv, xv.x, acc.x); acc.y = fmaf(wv, xv.y, acc.y);
          acc.z = fmaf(wv, xv.z, acc.z); acc.w = fmaf(wv, xv.w, acc.w);
          wsum += wv;
        }
        const int k = k0 + kk;
        if (RESIDENT) {
          float4* dst = reinterpret_cast<float4*>(acc_s + k * L + (l4 << 2));
          float4 cur = *dst;
          cur.x += acc.x; cur.y += acc.y; cur.z += acc.z; cur.w += acc.w;
          *dst = cur;
          if (l4 == 0) wsum_s[k] += wsum;
        } else if (SUMS) {
          float* dst = a.num + (int64_t)k * L + (l4 << 2);
          atomicAdd(dst + 0, acc.x); atomicAdd(dst + 1, acc.y);
          atomicAdd(dst + 2, acc.z); atomicAdd(dst + 3, acc.w);
          if (l4 == 0) atomicAdd(a.den + k, wsum);
        } else {
          const float4 m = *reinterpret_cast<const float4*>(mbase + kk * L + (l4 << 2));
          float* dst = a.gM + (int64_t)k * L + (l4 << 2);
          atomicAdd(dst + 0, -a.gscale * (acc.x - m.x * wsum));
          atomicAdd(dst + 1, -a.gscale * (acc.y - m.y * wsum));
          atomicAdd(dst + 2, -a.gscale * (acc.z - m.z * wsum));
          atomicAdd(dst + 3, -a.gscale * (acc.w - m.w * wsum));
        }
      }
      __syncthreads();
    }
    if (!SUMS && valid && a.gX != nullptr) {
#pragma unroll
      for (int c = 0; c < KHM_MAXCH; ++c) {
        if (NCH > 0 ? c < NCH : c < nch) {
          float4* dst = reinterpret_cast<float4*>(a.gX + i * a.ldg + ((c * TPP + s) << 2));
          float4 v = make_float4(a.gscale * g4[c].x, a.gscale * g4[c].y, a.gscale * g4[c].z, a.gscale * g4[c].w);
          if (a.accumulate_x) { const float4 o = *dst; v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w; }
          *dst = v;
        }
      }
    }
  }
  if (RESIDENT) {
    __syncthreads();
    for (int idx = threadIdx.x; idx < K * L; idx += blockDim.x) {
      const int k = idx / L;
      if (SUMS) atomicAdd(a.num + idx, acc_s[idx]);
      else atomicAdd(a.gM + idx, -a.gscale * (acc_s[idx] - ms[idx] * wsum_s[k]));
    }
    if (SUMS) for (int k = threadIdx.x; k < K; k += blockDim.x) atomicAdd(a.den + k, wsum_s[k]);
  }
  if (a.loss_sum != nullptr) {
    const double tot = block_sum<double>(lsum, red);
    if (threadIdx.x == 0) atomicAdd(a.loss_sum, tot);
  }
}

// ------------------------------------------------------------------------------------------
// pass 2, K <= 16 and L = 16*TPP (the HBM-side cases): two points per thread group, register-tiled tile product
// ------------------------------------------------------------------------------------------
// Same results as khm_pass2_kernel (same distances, harmonic sums and weights), about a third of its instructions:
//  * two points per thread: every broadcast LDS.128 of a centre chunk feeds both points (as in pass 1);
//  * four float4 chunks per lane (TPP = L/16 lanes per point): 64 registers of point data and gradient sums instead of
//    128, so four blocks fit per SM (the eight-chunk version ran 8 warps per SM at 240 registers and 32 % issue
//    utilisation: latency-bound);
//  * the point gradient sum_k w_k (x - m_k) is accumulated as (sum_k w_k) x - sum_k w_k m_k: one packed FMA per two
//    dimensions instead of a packed subtract + FMA (the distances keep the direct-difference form; here the terms that
//    cancel carry weights ~ d^2, so the L2 error of the gradient stays at fp32 rounding);
//  * the [K x pts] x [pts x L] product (centre gradient / centre sums) is register-tiled: a thread owns KT centres of one
//    float4 column for every PS-th point of the tile, its 4*KT accumulators live in registers for the whole kernel and the
//    weights are stored as duplicated pairs (w, w), so a 64-bit shared load IS the packed operand: per point and thread
//    one LDS.128 of x, ~KT/2 loads of weights and 2*KT FFMA2 (the scalar loop: 7 instructions per 4 FMA, 62 % of the
//    threads busy at K = 10, L = 64);
//  * ONE pass over the centres per tile: w_k = coef * u_k with coef = K / (e + eps)^2 known only after the harmonic sum,
//    so the loop accumulates sum_k u_k m_k from the centre chunk that is still in registers for the distance, the x tile
//    is stored pre-multiplied by coef and the point gradient is scaled at the end (no second sweep over the centres:
//    the shared-memory pipe was the busiest unit, 82 % of its wavefront rate);
//  * sum_i w_ik: per-point-slot partial sums in shared memory, added by the point's first lane;
//  * P4: p == 4 (the reference's Khp) known at compile time: no powf path in the code.
constexpr int FCH = 4;            // float4 chunks per lane in the fast kernel

template <int TPP>
__device__ __forceinline__ void load_point4(float4 (&x4)[FCH], const float* X, int64_t ldx, int64_t i, bool valid, int s) {
#pragma unroll
  for (int c = 0; c < FCH; ++c)
    x4[c] = valid ? *reinterpret_cast<const float4*>(X + i * ldx + ((c * TPP + s) << 2)) : make_float4(0.f, 0.f, 0.f, 0.f);
}

template <int TPP, int KT, bool SUMS, bool P4>
__global__ void __launch_bounds__(KHM_THREADS, 4) khm_pass2_fast_kernel(KhmArgs a, int KG) {
  extern __shared__ __align__(16) float smem[];
  const int pmode = P4 ? 4 : a.pmode;
  constexpr int PTS = KHM_THREADS / TPP;           // thread groups; a tile is 2*PTS points
  constexpr int L = 4 * FCH * TPP, L4 = L / 4;
  // row pitch of the x tile: a quarter-warp stores 8/TPP points x TPP float4 - the points' blocks must not share banks
  constexpr int XS = L + (TPP < 8 ? 4 * TPP : 4);
  constexpr int GS = (2 * KT + 3) & ~3;            // floats per centre group in a weight row (16-byte aligned groups)
  const int K = a.K;
  const int WS = KG * GS;                          // weight row: KG groups of KT duplicated pairs
  float* ms = smem;                                // [K][L] centres
  float* xs = ms + K * L;                          // [2*PTS][XS] x tile; [K][L] block sums at the end
  float* ws = xs + 2 * PTS * XS;                   // [2*PTS][WS]
  float* wsum = ws + 2 * PTS * WS;                 // [K] block sums of the weights (filled at the end)
  float* wpart = wsum + ((K + 3) & ~3);            // [K][PTS] per-point-slot partial sums of the weights
  __shared__ double red[32];
  const int tid = threadIdx.x, pt = tid / TPP, s = tid % TPP;
  const int64_t ntiles = (a.N + 2 * PTS - 1) / (2 * PTS);
  const float Kf = (float)K;
  double lsum = 0.0;
  stage_centres(ms, a.M, 0, K, L);
  for (int i = tid; i < 2 * PTS * WS; i += KHM_THREADS) ws[i] = 0.f;      // pairs of unused centre slots stay zero
  for (int i = tid; i < K * PTS; i += KHM_THREADS) wpart[i] = 0.f;
  // product role: column c4, centre group kg, point split ps
  const int c4 = tid % L4, kg = (tid / L4) % KG, ps = tid / (L4 * KG), PS = KHM_THREADS / (L4 * KG);
  f32x2 acc[KT][2];
#pragma unroll
  for (int j = 0; j < KT; ++j) acc[j][0] = acc[j][1] = 0ull;
  __syncthreads();
  for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const int64_t ia = t * 2 * PTS + pt, ib = ia + PTS;
    const bool va = ia < a.N, vb = ib < a.N;
    float4 xa[FCH], xb[FCH];
    load_point4<TPP>(xa, a.X, a.ldx, ia, va, s);
    load_point4<TPP>(xb, a.X, a.ldx, ib, vb, s);
    // ---- ONE pass over the centres: distance, harmonic sum and the un-normalised weight u_k (w_k = coef * u_k, coef
    // known only after the loop), whose products with the centre chunk accumulate while the chunk is still in registers
    __syncthreads();                                // the previous tile's product has read xs / ws
    float ea = 0.f, eb = 0.f, swa = 0.f, swb = 0.f;
    f32x2 ga[FCH][2], gb[FCH][2];
#pragma unroll
    for (int c = 0; c < FCH; ++c) ga[c][0] = ga[c][1] = gb[c][0] = gb[c][1] = 0ull;
    float* const wsa = ws + pt * WS;                // this point pair's weight rows
    float* const wsb = ws + (PTS + pt) * WS;
    for (int g = 0; g < KG; ++g) {
      float* const wga = wsa + g * GS;
      float* const wgb = wsb + g * GS;
#pragma unroll
      for (int j = 0; j < KT; ++j) {
        const int kk = g * KT + j;
        if (kk < K) {
          const float* mrow = ms + kk * L;
          f32x2 m0[FCH], m1[FCH];
          f32x2 a0 = 0ull, a1 = 0ull, b0 = 0ull, b1 = 0ull;
#pragma unroll
          for (int c = 0; c < FCH; ++c) {
            const float4 m = *reinterpret_cast<const float4*>(mrow + ((c * TPP + s) << 2));
            m0[c] = pk2(m.x, m.y); m1[c] = pk2(m.z, m.w);
            const f32x2 p0 = sub2(pk2(xa[c].x, xa[c].y), m0[c]), p1 = sub2(pk2(xa[c].z, xa[c].w), m1[c]);
            const f32x2 q0 = sub2(pk2(xb[c].x, xb[c].y), m0[c]), q1 = sub2(pk2(xb[c].z, xb[c].w), m1[c]);
            a0 = fma2(p0, p0, a0); a1 = fma2(p1, p1, a1);
            b0 = fma2(q0, q0, b0); b1 = fma2(q1, q1, b1);
          }
          float da, db;
          {
            float p, q, r, t;
            upk2(a0, p, q); upk2(a1, r, t);
            da = lanes_sum<TPP>((p + q) + (r + t));
            upk2(b0, p, q); upk2(b1, r, t);
            db = lanes_sum<TPP>((p + q) + (r + t));
          }
          const float pa = pow_p(da, a.p, pmode), pb = pow_p(db, a.p, pmode);
          const float ra = rcp_fast(pa + KHM_EPS), rb = rcp_fast(pb + KHM_EPS);
          ea += ra; eb += rb;
          float ua, ub;
          if (SUMS) {
            ua = rcp_fast(pa * da + KHM_EPS);                       // 1 / (d^(p+2) + eps)
            ub = rcp_fast(pb * db + KHM_EPS);
          } else {
            ua = da > 0.f ? a.p * pow_pm2(da, a.p, pmode) * (ra * ra) : 0.f;   // p d^(p-2) / (d^p + eps)^2
            ub = db > 0.f ? a.p * pow_pm2(db, a.p, pmode) * (rb * rb) : 0.f;
          }
          if (!SUMS) {
            const f32x2 ua2 = pk2(ua, ua), ub2 = pk2(ub, ub);
#pragma unroll
            for (int c = 0; c < FCH; ++c) {
              fma2_acc(ga[c][0], ua2, m0[c]); fma2_acc(ga[c][1], ua2, m1[c]);
              fma2_acc(gb[c][0], ub2, m0[c]); fma2_acc(gb[c][1], ub2, m1[c]);
            }
            swa += ua; swb += ub;
          }
          if (s == 0) {
            *reinterpret_cast<float2*>(wga + 2 * j) = make_float2(ua, ua);
            *reinterpret_cast<float2*>(wgb + 2 * j) = make_float2(ub, ub);
          }
        }
      }
    }
    if (s == 0) {
      if (va) lsum += (double)(Kf / (ea + KHM_EPS));
      if (vb) lsum += (double)(Kf / (eb + KHM_EPS));
    }
    // per-point coefficient: gradient K/(e+eps)^2 ; centre update alpha = 1/(e^2+eps); zero for the padding points
    float ca = SUMS ? 1.0f / (ea * ea + KHM_EPS) : Kf / ((ea + KHM_EPS) * (ea + KHM_EPS));
    float cb = SUMS ? 1.0f / (eb * eb + KHM_EPS) : Kf / ((eb + KHM_EPS) * (eb + KHM_EPS));
    if (!va) ca = 0.f;
    if (!vb) cb = 0.f;
    // the x tile carries the coefficient: sum_i w_ik x_i = sum_i u_ik (coef_i x_i)
    {
      const f32x2 ca2 = pk2(ca, ca), cb2 = pk2(cb, cb);
#pragma unroll
      for (int c = 0; c < FCH; ++c) {
        float4 v, u;
        upk2(mul2(ca2, pk2(xa[c].x, xa[c].y)), v.x, v.y); upk2(mul2(ca2, pk2(xa[c].z, xa[c].w)), v.z, v.w);
        upk2(mul2(cb2, pk2(xb[c].x, xb[c].y)), u.x, u.y); upk2(mul2(cb2, pk2(xb[c].z, xb[c].w)), u.z, u.w);
        *reinterpret_cast<float4*>(xs + pt * XS + ((c * TPP + s) << 2)) = v;
        *reinterpret_cast<float4*>(xs + (PTS + pt) * XS + ((c * TPP + s) << 2)) = u;
      }
    }
    // sum_i w_ik: the TPP lanes of a point share its K centres (lane s takes k = s, s + TPP, ...) and add coef * u_k of
    // the two points to the slot's partial sums (the u_k pairs were written by the point's first lane)
    __syncwarp();
    for (int kk = s; kk < K; kk += TPP) {
      const int o = (kk / KT) * GS + 2 * (kk % KT);
      wpart[kk * PTS + pt] += ca * wsa[o] + cb * wsb[o];
    }
    if (!SUMS && a.gX != nullptr) {
      const float fa = ca * a.gscale, fb = cb * a.gscale;
      const f32x2 sa2 = pk2(swa * fa, swa * fa), sb2 = pk2(swb * fb, swb * fb);
      const f32x2 na = pk2(-fa, -fa), nb = pk2(-fb, -fb);
#pragma unroll
      for (int c = 0; c < FCH; ++c) {
        // gscale * coef * (sum_k u_k * x - sum_k u_k m_k)
        float4 v, u;
        upk2(fma2(sa2, pk2(xa[c].x, xa[c].y), mul2(na, ga[c][0])), v.x, v.y);
        upk2(fma2(sa2, pk2(xa[c].z, xa[c].w), mul2(na, ga[c][1])), v.z, v.w);
        upk2(fma2(sb2, pk2(xb[c].x, xb[c].y), mul2(nb, gb[c][0])), u.x, u.y);
        upk2(fma2(sb2, pk2(xb[c].z, xb[c].w), mul2(nb, gb[c][1])), u.z, u.w);
        if (va) {
          float4* dst = reinterpret_cast<float4*>(a.gX + ia * a.ldg + ((c * TPP + s) << 2));
          if (a.accumulate_x) { const float4 o = *dst; v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w; }
          *dst = v;
        }
        if (vb) {
          float4* dst = reinterpret_cast<float4*>(a.gX + ib * a.ldg + ((c * TPP + s) << 2));
          if (a.accumulate_x) { const float4 o = *dst; u.x += o.x; u.y += o.y; u.z += o.z; u.w += o.w; }
          *dst = u;
        }
      }
    }
    __syncthreads();
    // ---- tile product: acc[j] += w[q][kg*KT + j] * x[q][4*c4 .. 4*c4+3] over the thread's points
    {
      const float* xq = xs + (c4 << 2) + ps * XS;
      const float* wq = ws + kg * GS + ps * WS;
      const int xstep = PS * XS, wstep = PS * WS;
#pragma unroll 4
      for (int q = ps; q < 2 * PTS; q += PS, xq += xstep, wq += wstep) {
        const float4 xv = *reinterpret_cast<const float4*>(xq);
        const f32x2 x0 = pk2(xv.x, xv.y), x1 = pk2(xv.z, xv.w);
        f32x2 w2[KT];
#pragma unroll
        for (int j = 0; j + 1 < KT; j += 2) {
          const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(wq + 2 * j);
          w2[j] = v.x; w2[j + 1] = v.y;
        }
        if (KT & 1) w2[KT - 1] = *reinterpret_cast<const f32x2*>(wq + 2 * (KT - 1));
#pragma unroll
        for (int j = 0; j < KT; ++j) {
          acc[j][0] = fma2(w2[j], x0, acc[j][0]);
          acc[j][1] = fma2(w2[j], x1, acc[j][1]);
        }
      }
    }
  }
  // ---- block sums -> global
  __syncthreads();
  float* acc_s = xs;
  for (int i = tid; i < K * L; i += KHM_THREADS) acc_s[i] = 0.f;
  __syncthreads();
#pragma unroll
  for (int j = 0; j < KT; ++j) {
    const int k = kg * KT + j;
    if (k < K) {
      float p0, p1, p2, p3;
      upk2(acc[j][0], p0, p1); upk2(acc[j][1], p2, p3);
      float* dst = acc_s + k * L + (c4 << 2);
      atomicAdd(dst + 0, p0); atomicAdd(dst + 1, p1); atomicAdd(dst + 2, p2); atomicAdd(dst + 3, p3);
    }
  }
  // sum_i w_ik of the block
  for (int k = tid >> 5; k < K; k += KHM_THREADS / 32) {
    float v = 0.f;
    for (int i = tid & 31; i < PTS; i += 32) v += wpart[k * PTS + i];
    v = warp_sum(v);
    if ((tid & 31) == 0) wsum[k] = v;
  }
  __syncthreads();
  for (int idx = tid; idx < K * L; idx += KHM_THREADS) {
    const int k = idx / L;
    if (SUMS) atomicAdd(a.num + idx, acc_s[idx]);
    else atomicAdd(a.gM + idx, -a.gscale * (acc_s[idx] - ms[idx] * wsum[k]));
  }
  if (SUMS) for (int k = tid; k < K; k += KHM_THREADS) atomicAdd(a.den + k, wsum[k]);
  if (a.loss_sum != nullptr) {
    const double tot = block_sum<double>(lsum, red);
    if (tid == 0) atomicAdd(a.loss_sum, tot);
  }
}

__global__ void group_argmin_kernel(const float* dist, int64_t G, int K, int32_t* gid) {
  const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= G) return;
  float best = dist[g * K];
  int bi = 0;
  for (int k = 1; k < K; ++k) {
    const float v = dist[g * K + k];
    if (v < best) { best = v; bi = k; }
  }
  gid[g] = bi;
}

__global__ void center_apply_kernel(const float* num, const float* den, float* M, int K, int L) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < K * L) M[idx] = num[idx] / den[idx / L];
}

constexpr int RESIDENT_SMEM_LIMIT = 96 * 1024;

int pick_tpp(int L) {
  if (L <= 0 || (L & 3)) return 0;
  for (int tpp = 1; tpp <= 32; tpp <<= 1)
    if (L % (4 * tpp) == 0 && L / (4 * tpp) <= KHM_MAXCH) return tpp;
  return 0;
}

int check_common(const char* name, const float* X, int64_t ldx, const float* M, int64_t N, int K, int L) {
  LSHM_REQUIRE(X && M, "%s: null pointer", name);
  LSHM_REQUIRE(N >= 0 && K > 0 && L > 0, "%s: bad sizes N=%lld K=%d L=%d", name, (long long)N, K, L);
  LSHM_REQUIRE(pick_tpp(L) != 0, "%s: latent dim L=%d unsupported (need L%%4==0 and L<=1024 with L/(4*2^j)<=8)", name, L);
  LSHM_REQUIRE(ldx >= L && (ldx & 3) == 0, "%s: row stride %lld must be >=L and a multiple of 4", name, (long long)ldx);
  LSHM_REQUIRE((reinterpret_cast<uintptr_t>(X) & 15) == 0 && (reinterpret_cast<uintptr_t>(M) & 15) == 0,
               "%s: X and M must be 16-byte aligned", name);
  return LSHM_OK;
}

int pmode_of(float p) { return p == 2.f ? 2 : (p == 4.f ? 4 : 0); }

template <int TPP, bool RES, int NCH, int MODE>
int launch_pass1_m(const KhmArgs& a, size_t smem, int grid, cudaStream_t st) {
  if (smem > 48 * 1024)
    LSHM_CUDA(cudaFuncSetAttribute(khm_pass1_kernel<TPP, RES, NCH, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "khm_pass1");
  khm_pass1_kernel<TPP, RES, NCH, MODE><<<grid, KHM_THREADS, smem, st>>>(a);
  return LSHM_OK;
}

template <int TPP, bool RES, int NCH>
int launch_pass1_n(const KhmArgs& a, size_t smem, int grid, cudaStream_t st) {
  int rc;
  if (a.dist != nullptr || (a.ids != nullptr && (a.loss_sum != nullptr || a.e_out != nullptr)))
    rc = launch_pass1_m<TPP, RES, NCH, 2>(a, smem, grid, st);
  else if (a.ids != nullptr)
    rc = launch_pass1_m<TPP, RES, NCH, 1>(a, smem, grid, st);
  else
    rc = launch_pass1_m<TPP, RES, NCH, 0>(a, smem, grid, st);
  if (rc) return rc;
  LSHM_CHECK_LAUNCH("khm_pass1");
  return LSHM_OK;
}

template <int TPP, bool RES>
int launch_pass1_t(const KhmArgs& a, size_t smem, int grid, cudaStream_t st) {
  if (a.L == 4 * TPP * KHM_MAXCH) return launch_pass1_n<TPP, RES, KHM_MAXCH>(a, smem, grid, st);
  if (a.L == 4 * TPP * 4) return launch_pass1_n<TPP, RES, 4>(a, smem, grid, st);
  return launch_pass1_n<TPP, RES, 0>(a, smem, grid, st);
}

template <int TPP, bool RES, bool SUMS, int NCH>
int launch_pass2_n(const KhmArgs& a, size_t smem, int grid, cudaStream_t st) {
  if (smem > 48 * 1024)
    LSHM_CUDA(cudaFuncSetAttribute(khm_pass2_kernel<TPP, RES, SUMS, NCH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "khm_pass2");
  khm_pass2_kernel<TPP, RES, SUMS, NCH><<<grid, KHM_THREADS, smem, st>>>(a);
  LSHM_CHECK_LAUNCH("khm_pass2");
  return LSHM_OK;
}

template <int TPP, bool RES, bool SUMS>
int launch_pass2_t(const KhmArgs& a, size_t smem, int grid, cudaStream_t st) {
  return a.L == 4 * TPP * KHM_MAXCH ? launch_pass2_n<TPP, RES, SUMS, KHM_MAXCH>(a, smem, grid, st)
                                    : launch_pass2_n<TPP, RES, SUMS, 0>(a, smem, grid, st);
}

int grid_for(int64_t N, int tpp, int blocks_per_sm, int ppt = 1) {
  const int pts = ppt * KHM_THREADS / tpp;
  const int64_t ntiles = ceil_div(N, pts);
  const int64_t cap = (int64_t)sm_count() * blocks_per_sm;
  return (int)(ntiles < cap ? (ntiles > 0 ? ntiles : 1) : cap);
}

int launch_pass1(const KhmArgs& a, cudaStream_t st) {
  const int tpp = pick_tpp(a.L);
  const size_t res_bytes = (size_t)a.K * a.L * sizeof(float);
  const bool res = res_bytes <= RESIDENT_SMEM_LIMIT;
  size_t smem = res ? res_bytes : (size_t)KHM_KC * a.L * sizeof(float);
  if (tpp == 1 && a.L == 32) smem += (size_t)(KHM_THREADS / 32) * 32 * 36 * sizeof(float);   // per-warp transpose buffers
  const int grid = grid_for(a.N, tpp, 8, 2);
#define P1(T) (res ? launch_pass1_t<T, true>(a, smem, grid, st) : launch_pass1_t<T, false>(a, smem, grid, st))
  switch (tpp) {
    case 1: return P1(1);
    case 2: return P1(2);
    case 4: return P1(4);
    case 8: return P1(8);
    case 16: return P1(16);
    default: return P1(32);
  }
#undef P1
}

// (KT, KG) of the register-tiled product: KG groups of KT centres cover K; 128 threads = L/4 columns x KG x point splits.
// The fast kernel holds FCH = 4 chunks per lane: tpp = L/16 lanes per point.
bool pick_fast(int K, int L, int* tpp, int* kt, int* kg) {
  if (K > 16 || L < 32 || L > 256 || (L & (L - 1))) return false;
  *tpp = L / 16;
  for (int t : {4, 5, 8}) {
    const int g = (K + t - 1) / t;
    if ((g == 1 || g == 2 || g == 4) && (L / 4) * g <= KHM_THREADS) { *kt = t; *kg = g; return true; }
  }
  return false;
}

template <int TPP, int KT, bool SUMS, bool P4>
int launch_pass2_fast_p(const KhmArgs& a, int kg, cudaStream_t st) {
  const int pts = KHM_THREADS / TPP, gs = (2 * KT + 3) & ~3;
  const int xs = a.L + (TPP < 8 ? 4 * TPP : 4);
  const size_t smem = ((size_t)a.K * a.L + (size_t)2 * pts * xs + (size_t)2 * pts * kg * gs + (size_t)((a.K + 3) & ~3) +
                       (size_t)a.K * pts) * sizeof(float);
  if (smem > 48 * 1024)
    LSHM_CUDA(cudaFuncSetAttribute(khm_pass2_fast_kernel<TPP, KT, SUMS, P4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "khm_pass2");
  int per_sm = 4;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, khm_pass2_fast_kernel<TPP, KT, SUMS, P4>, KHM_THREADS, smem);
  const int64_t ntiles = ceil_div(a.N, 2 * pts);
  const int64_t cap = (int64_t)sm_count() * std::max(1, per_sm);
  const int grid = (int)std::max<int64_t>(1, std::min(ntiles, cap));
  khm_pass2_fast_kernel<TPP, KT, SUMS, P4><<<grid, KHM_THREADS, smem, st>>>(a, kg);
  LSHM_CHECK_LAUNCH("khm_pass2");
  return LSHM_OK;
}

template <int TPP, int KT, bool SUMS>
int launch_pass2_fast_k(const KhmArgs& a, int kg, cudaStream_t st) {
  return a.pmode == 4 ? launch_pass2_fast_p<TPP, KT, SUMS, true>(a, kg, st) : launch_pass2_fast_p<TPP, KT, SUMS, false>(a, kg, st);
}

template <int TPP, bool SUMS>
int launch_pass2_fast(const KhmArgs& a, int kt, int kg, cudaStream_t st) {
  switch (kt) {
    case 4: return launch_pass2_fast_k<TPP, 4, SUMS>(a, kg, st);
    case 5: return launch_pass2_fast_k<TPP, 5, SUMS>(a, kg, st);
    default: return launch_pass2_fast_k<TPP, 8, SUMS>(a, kg, st);
  }
}

template <bool SUMS>
int launch_pass2(const KhmArgs& a, cudaStream_t st) {
  const int tpp = pick_tpp(a.L);
  int kt = 0, kg = 0, ftpp = 0;
  static const bool no_fast = getenv("LSHM_KHM_NOFAST") != nullptr;       // experiment switch
  if (!no_fast && pick_fast(a.K, a.L, &ftpp, &kt, &kg)) {
    switch (ftpp) {
      case 2: return launch_pass2_fast<2, SUMS>(a, kt, kg, st);
      case 4: return launch_pass2_fast<4, SUMS>(a, kt, kg, st);
      case 8: return launch_pass2_fast<8, SUMS>(a, kt, kg, st);
      default: return launch_pass2_fast<16, SUMS>(a, kt, kg, st);
    }
  }
  const int pts = KHM_THREADS / tpp;
  const size_t tile = ((size_t)pts * a.L + (size_t)pts * KHM_KC + (a.K <= KHM_KC ? (size_t)pts * KHM_KC : 0)) * sizeof(float);
  const size_t res_bytes = ((size_t)2 * a.K * a.L + ((a.K + 3) & ~3)) * sizeof(float) + tile;
  const bool res = res_bytes <= RESIDENT_SMEM_LIMIT;
  const size_t smem = res ? res_bytes : (size_t)KHM_KC * a.L * sizeof(float) + tile;
  const int grid = grid_for(a.N, tpp, 4);
#define P2(T) (res ? launch_pass2_t<T, true, SUMS>(a, smem, grid, st) : launch_pass2_t<T, false, SUMS>(a, smem, grid, st))
  switch (tpp) {
    case 1: return P2(1);
    case 2: return P2(2);
    case 4: return P2(4);
    case 8: return P2(8);
    case 16: return P2(16);
    default: return P2(32);
  }
#undef P2
}

}  // namespace
}  // namespace lshm

using namespace lshm;

extern "C" {

int lshm_khm_fwd(const float* X, int64_t ldx, const float* M, int64_t N, int K, int L, float p,
                 double* loss_sum, float* e_out, lshm_stream_t stream) {
  if (int rc = check_common("lshm_khm_fwd", X, ldx, M, N, K, L)) return rc;
  LSHM_REQUIRE(loss_sum != nullptr, "lshm_khm_fwd: loss_sum is null");
  if (N == 0) return LSHM_OK;
  KhmArgs a{};
  a.X = X; a.ldx = ldx; a.M = M; a.N = N; a.K = K; a.L = L; a.p = p; a.pmode = pmode_of(p);
  a.loss_sum = loss_sum; a.e_out = e_out;
  return launch_pass1(a, as_stream(stream));
}

int lshm_khm_assign(const float* X, int64_t ldx, const float* M, int64_t N, int K, int L,
                    int32_t* ids, lshm_stream_t stream) {
  if (int rc = check_common("lshm_khm_assign", X, ldx, M, N, K, L)) return rc;
  LSHM_REQUIRE(ids != nullptr, "lshm_khm_assign: ids is null");
  if (N == 0) return LSHM_OK;
  KhmArgs a{};
  a.X = X; a.ldx = ldx; a.M = M; a.N = N; a.K = K; a.L = L; a.p = 2.f; a.pmode = 2;
  a.ids = ids;
  return launch_pass1(a, as_stream(stream));
}

int lshm_khm_group_dist(const float* X, int64_t ldx, const float* M, int64_t N, int K, int L,
                        float p, int group, float* dist, int32_t* gid, lshm_stream_t stream) {
  if (int rc = check_common("lshm_khm_group_dist", X, ldx, M, N, K, L)) return rc;
  LSHM_REQUIRE(dist && gid, "lshm_khm_group_dist: null output");
  LSHM_REQUIRE(group > 0 && N % group == 0, "lshm_khm_group_dist: N=%lld not a multiple of group=%d", (long long)N, group);
  if (N == 0) return LSHM_OK;
  const int64_t G = N / group;
  cudaStream_t st = as_stream(stream);
  LSHM_CUDA(cudaMemsetAsync(dist, 0, sizeof(float) * G * K, st), "lshm_khm_group_dist");
  KhmArgs a{};
  a.X = X; a.ldx = ldx; a.M = M; a.N = N; a.K = K; a.L = L; a.p = p; a.pmode = pmode_of(p);
  a.dist = dist; a.group = group;
  if (int rc = launch_pass1(a, st)) return rc;
  group_argmin_kernel<<<(unsigned)ceil_div(G, 128), 128, 0, st>>>(dist, G, K, gid);
  LSHM_CHECK_LAUNCH("lshm_khm_group_dist");
  return LSHM_OK;
}

static int khm_bwd_common(const char* name, const float* X, int64_t ldx, const float* M, int64_t N,
                          int K, int L, float p, float gscale, double* loss_sum, float* gX,
                          int64_t ldg, int accumulate_x, float* gM, lshm_stream_t stream) {
  if (int rc = check_common(name, X, ldx, M, N, K, L)) return rc;
  LSHM_REQUIRE(gM != nullptr, "%s: gM is null", name);
  if (gX) {
    LSHM_REQUIRE(ldg >= L && (ldg & 3) == 0 && (reinterpret_cast<uintptr_t>(gX) & 15) == 0,
                 "%s: gX must be 16-byte aligned with row stride %%4==0", name);
  }
  if (N == 0) return LSHM_OK;
  KhmArgs a{};
  a.X = X; a.ldx = ldx; a.M = M; a.N = N; a.K = K; a.L = L; a.p = p; a.pmode = pmode_of(p);
  a.loss_sum = loss_sum; a.gscale = gscale; a.gX = gX; a.ldg = ldg; a.accumulate_x = accumulate_x; a.gM = gM;
  return launch_pass2<false>(a, as_stream(stream));
}

int lshm_khm_bwd(const float* X, int64_t ldx, const float* M, int64_t N, int K, int L, float p,
                 float gscale, float* gX, int64_t ldg, int accumulate_x, float* gM,
                 lshm_stream_t stream) {
  return khm_bwd_common("lshm_khm_bwd", X, ldx, M, N, K, L, p, gscale, nullptr, gX, ldg, accumulate_x, gM, stream);
}

int lshm_khm_fwd_bwd(const float* X, int64_t ldx, const float* M, int64_t N, int K, int L,
                     float p, float gscale, double* loss_sum, float* gX, int64_t ldg,
                     int accumulate_x, float* gM, lshm_stream_t stream) {
  LSHM_REQUIRE(loss_sum != nullptr, "lshm_khm_fwd_bwd: loss_sum is null");
  return khm_bwd_common("lshm_khm_fwd_bwd", X, ldx, M, N, K, L, p, gscale, loss_sum, gX, ldg, accumulate_x, gM, stream);
}

int lshm_khm_center_sums(const float* X, int64_t ldx, const float* M, int64_t N, int K, int L,
                         float p, float* num, float* den, lshm_stream_t stream) {
  if (int rc = check_common("lshm_khm_center_sums", X, ldx, M, N, K, L)) return rc;
  LSHM_REQUIRE(num && den, "lshm_khm_center_sums: null output");
  if (N == 0) return LSHM_OK;
  KhmArgs a{};
  a.X = X; a.ldx = ldx; a.M = M; a.N = N; a.K = K; a.L = L; a.p = p; a.pmode = pmode_of(p);
  a.num = num; a.den = den;
  return launch_pass2<true>(a, as_stream(stream));
}

int lshm_khm_center_apply(const float* num, const float* den, float* M, int K, int L,
                          lshm_stream_t stream) {
  LSHM_REQUIRE(num && den && M && K > 0 && L > 0, "lshm_khm_center_apply: bad arguments");
  center_apply_kernel<<<(unsigned)ceil_div((int64_t)K * L, 256), 256, 0, as_stream(stream)>>>(num, den, M, K, L);
  LSHM_CHECK_LAUNCH("lshm_khm_center_apply");
  return LSHM_OK;
}

}  // extern "C"
