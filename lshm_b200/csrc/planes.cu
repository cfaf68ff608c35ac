// Operand planes (see tma.cuh): generic fp32 -> planes staging kernels, the fused writers of the training
// step (residual split -> 1-D planes, gradient combine -> 2-D planes) and the host-side tensor-map factory.
//
// Reference semantics of what is being staged: the inputs of Conv2d(k4,s2,p1) / Conv1d(k4,s4,p1) and the
// output gradients of ConvTranspose2d / ConvTranspose1d, /root/reference/src/lofar_models.py:73,93,158,178;
// x11 = (x - x1)/2 and its transpose, /root/reference/src/kharmonic_lofar.py:137-144.
#include <mutex>
#include "conv_geom.cuh"
#include "tma.cuh"

namespace lshm {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) == cudaSuccess &&
        qr == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

int make_plane_tmap(CUtensorMap* m, const void* half_base, int64_t Qs, int chunks, int slots, int box_chunks) {
  EncodeTiledFn fn = encode_tiled_fn();
  LSHM_REQUIRE(fn != nullptr, "operand planes: cuTensorMapEncodeTiled is not available from this driver");
  LSHM_REQUIRE(slots >= PLANE_ROW && slots % PLANE_ROW == 0 && slots / PLANE_ROW <= 256 && box_chunks >= 1 && box_chunks <= 256 &&
               Qs % PLANE_ROW == 0, "operand planes: bad box");
  LSHM_REQUIRE((reinterpret_cast<uintptr_t>(half_base) & 15) == 0, "operand planes: buffer must be 16-byte aligned");
  const cuuint64_t gdim[3] = {64, (cuuint64_t)(Qs / PLANE_ROW), (cuuint64_t)chunks};
  const cuuint64_t gstr[2] = {512, (cuuint64_t)Qs * 16};          // bytes, dims 1 and 2
  const cuuint32_t box[3] = {64, (cuuint32_t)(slots / PLANE_ROW), (cuuint32_t)box_chunks};
  const cuuint32_t estr[3] = {1, 1, 1};
  const CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_UINT64, 3, const_cast<void*>(half_base), gdim, gstr, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  LSHM_REQUIRE(r == CUDA_SUCCESS, "operand planes: cuTensorMapEncodeTiled failed (%d)", (int)r);
  return LSHM_OK;
}

namespace {

using namespace tc;

// Q here is the chunk stride in positions (PlaneGeom::Qs)
__device__ __forceinline__ void store_chunk(uint8_t* __restrict__ hi, size_t half_bytes, int64_t Q, int cc, int64_t q,
                                            const float (&v)[8]) {
  uint4 h, l;
  split8(v, h, l);
  uint8_t* p = hi + ((size_t)cc * (size_t)Q + (size_t)q) * 16;
  *reinterpret_cast<uint4*>(p) = h;
  *reinterpret_cast<uint4*>(p + half_bytes) = l;
}

// one thread per (chunk, block position): 2 channels x the 2x2 pixel block at rows 2by-1, 2by / columns 2bx-1, 2bx
__global__ void __launch_bounds__(256)
stage2d_kernel(const float* __restrict__ big, int64_t big_ns, uint8_t* __restrict__ planes, size_t half_bytes,
               int Bc, int h, int w, int64_t Q, int64_t Qs, int chunks, FastDiv d_pp, FastDiv d_pw) {
  const int PW = w + 1, PH = h + 1, W = 2 * w;
  const int64_t HW = 4 * (int64_t)h * w;
  const int64_t total = Q * chunks;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int cc = (int)(idx / Q);
    const uint32_t q = (uint32_t)(idx - (int64_t)cc * Q);
    const uint32_t n = fdiv(q, d_pp), r = q - n * (uint32_t)(PH * PW);
    const int by = (int)fdiv(r, d_pw), bx = (int)r - by * PW;
    const bool r0 = by > 0, r1 = by < h, c0 = bx > 0, c1 = bx < w;
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = 0.f;
#pragma unroll
    for (int bb = 0; bb < 2; ++bb) {
      const int b = 2 * cc + bb;
      if (b < Bc) {
        const float* p = big + (int64_t)n * big_ns + (int64_t)b * HW + (int64_t)(2 * by - 1) * W + (2 * bx - 1);
        if (r0 && c0) v[bb * 4 + 0] = __ldg(p);
        if (r0 && c1) v[bb * 4 + 1] = __ldg(p + 1);
        if (r1 && c0) v[bb * 4 + 2] = __ldg(p + W);
        if (r1 && c1) v[bb * 4 + 3] = __ldg(p + W + 1);
      }
    }
    store_chunk(planes, half_bytes, Qs, cc, q, v);
  }
}

// one thread per (chunk, window j): 2 channels x samples [4j - pad, 4j - pad + 3]
__global__ void __launch_bounds__(256)
stage1d_kernel(const float* __restrict__ big, int64_t big_ns, uint8_t* __restrict__ planes, size_t half_bytes,
               int Bc, int l, int pad, int64_t Q, int64_t Qs, int chunks, FastDiv d_l) {
  const int64_t total = Q * chunks;
  const int64_t Lb = 4 * (int64_t)l;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int cc = (int)(idx / Q);
    const uint32_t q = (uint32_t)(idx - (int64_t)cc * Q);
    const uint32_t n = fdiv(q, d_l);
    const int j = (int)(q - n * (uint32_t)l);
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = 0.f;
#pragma unroll
    for (int bb = 0; bb < 2; ++bb) {
      const int b = 2 * cc + bb;
      if (b < Bc) {
        const float* p = big + (int64_t)n * big_ns + (int64_t)b * Lb + 4 * (int64_t)j - pad;
#pragma unroll
        for (int t = 0; t < 4; ++t)
          if (pad == 0 || t > 0 || j > 0) v[bb * 4 + t] = __ldg(p + t);
      }
    }
    store_chunk(planes, half_bytes, Qs, cc, q, v);
  }
}

// ------------------------------------------------------------------------------------------------
// Fused writers.  Both work on 32 x 32 pixel tiles of a PAIR of channel planes (a chunk holds two channels).
constexpr int PT = 32;     // P must be a multiple of the 32-row strips

// ---- micro-tile scheme of the fused writers ------------------------------------------------------------
// A thread owns a 4 x 4 (t x f) block of a PAIR of channel planes; a warp covers 32 t x 16 f (lane = 4*tg + fg:
// 8 row groups x 4 column groups), a 256-thread block a strip of 32 rows x 128 columns.  Row-major tensors are
// read as float4 along f (a warp touches 64 contiguous bytes in each of 8 rows), transposed tensors as float4
// along t (128 contiguous bytes per column): every 32-byte sector is used whole in both layouts, no shared
// memory, no barriers.  A chunk of the time-axis planes is 4 consecutive f of one row, a chunk of the
// frequency-axis planes 4 consecutive t of one column - both are rows / columns of the thread's own block.

// x11 = (x - x1) / 2 written as the 1-D pad-1 planes of the time-axis net (flattened s = t*P + f) and of the
// frequency-axis net (flattened s = f*P + t): window j of a flattened map covers samples [4j-1, 4j+2], i.e. the
// element BEFORE the block's row / column and its first three elements (one halo row + one halo column per thread;
// the flattened predecessor of (t, f=0) is (t-1, P-1), of (f, t=0) it is (f-1, P-1)).
#ifndef LSHM_SPLIT_MINB
#define LSHM_SPLIT_MINB 3
#endif
__global__ void __launch_bounds__(256, LSHM_SPLIT_MINB)
residual_split_planes_kernel(const float* __restrict__ x, const float* __restrict__ x1, uint8_t* __restrict__ pT,
                             uint8_t* __restrict__ pF, size_t half_bytes, int C, int P, int64_t Qs) {
  const int strips = P / 32;
  const int strip = blockIdx.x % strips;
  const int64_t pair = blockIdx.x / strips;
  const int ccn = C / 2;
  const int n = (int)(pair / ccn), cc = (int)(pair % ccn);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int t4 = strip * 32 + (lane >> 2) * 4;
  const int f4 = warp * 16 + (lane & 3) * 4;
  if (f4 >= P) return;
  float v[2][4][4], left[2][4], up[2][4];
#pragma unroll
  for (int ch = 0; ch < 2; ++ch) {
    const int64_t plane = ((int64_t)n * C + 2 * cc + ch) * P * (int64_t)P;
    const float* xp = x + plane;
    const float* x1p = x1 + plane;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t off = (int64_t)(t4 + j) * P + f4;
      const float4 a = __ldg(reinterpret_cast<const float4*>(xp + off));
      const float4 b = __ldg(reinterpret_cast<const float4*>(x1p + off));
      v[ch][j][0] = 0.5f * (a.x - b.x); v[ch][j][1] = 0.5f * (a.y - b.y);
      v[ch][j][2] = 0.5f * (a.z - b.z); v[ch][j][3] = 0.5f * (a.w - b.w);
      // element before (t, f4) in the flattened time-axis map
      const int64_t lo = f4 > 0 ? off - 1 : off - 1;       // (t, f4-1), or (t-1, P-1) when f4 == 0: both are off-1
      left[ch][j] = (f4 > 0 || t4 + j > 0) ? 0.5f * (__ldg(xp + lo) - __ldg(x1p + lo)) : 0.f;
    }
    // elements before (f, t4) in the flattened frequency-axis map: (t4-1, f), or (P-1, f-1) when t4 == 0
    if (t4 > 0) {
      const int64_t off = (int64_t)(t4 - 1) * P + f4;
      const float4 a = __ldg(reinterpret_cast<const float4*>(xp + off));
      const float4 b = __ldg(reinterpret_cast<const float4*>(x1p + off));
      up[ch][0] = 0.5f * (a.x - b.x); up[ch][1] = 0.5f * (a.y - b.y);
      up[ch][2] = 0.5f * (a.z - b.z); up[ch][3] = 0.5f * (a.w - b.w);
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int f = f4 + k;
        const int64_t off = (int64_t)(P - 1) * P + (f - 1);
        up[ch][k] = f > 0 ? 0.5f * (__ldg(xp + off) - __ldg(x1p + off)) : 0.f;
      }
    }
  }
  const int64_t l = (int64_t)P * P / 4;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float c[8] = {left[0][j], v[0][j][0], v[0][j][1], v[0][j][2], left[1][j], v[1][j][0], v[1][j][1], v[1][j][2]};
    store_chunk(pT, half_bytes, Qs, cc, (int64_t)n * l + ((int64_t)(t4 + j) * P + f4) / 4, c);
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float c[8] = {up[0][k], v[0][0][k], v[0][1][k], v[0][2][k], up[1][k], v[1][0][k], v[1][1][k], v[1][2][k]};
    store_chunk(pF, half_bytes, Qs, cc, (int64_t)n * l + ((int64_t)(f4 + k) * P + t4) / 4, c);
  }
}

// gx1 = g1p - 0.5 (gT + transpose(gF)) written as the 2-D planes of the 2-D net's last transposed conv
// (block (by,bx) = pixel rows 2by-1, 2by x columns 2bx-1, 2bx), plus the per-channel sums (bias gradient).
// Micro-tile scheme with the tile SHIFTED by (-1,-1): the thread of (t4, f4) evaluates gx1 on rows t4-1 .. t4+2 x
// columns f4-1 .. f4+2 (row-major inputs: one scalar + one float4 per row; the transposed input gF: one scalar + one
// float4 along t per column) = the four 2 x 2 pixel blocks by = t4/2, t4/2+1, bx = f4/2, f4/2+1.  Threads on the
// bottom / right edge also emit the last halo block row / column (pixel row / column P-1 and the zero beyond it).
__device__ __forceinline__ float gx1_at(const float* __restrict__ ap, const float* __restrict__ bp,
                                        const float* __restrict__ cp, int P, int t, int f) {
  const int64_t off = (int64_t)t * P + f;
  return __ldg(ap + off) - 0.5f * (__ldg(bp + off) + __ldg(cp + (int64_t)f * P + t));
}

#ifndef LSHM_COMBINE_MINB
#define LSHM_COMBINE_MINB 4
#endif
__global__ void __launch_bounds__(256, LSHM_COMBINE_MINB)
combine_planes_kernel(const float* __restrict__ g1p, const float* __restrict__ gT, const float* __restrict__ gF,
                      uint8_t* __restrict__ planes, size_t half_bytes, int C, int P, int64_t Qs, int64_t items,
                      float* __restrict__ db1) {
  __shared__ float cacc[64];                 // per-channel sums of this (persistent) block
  // the 4 chunks of a thread are 16 bytes at a 32-byte stride in the planes: they go through this per-warp staging
  // area and leave as 128-byte runs (8 chunks of one block row)
  __shared__ uint4 stg[8][2][16][8];         // [warp][half][block row 2 tg + bi][block column 2 fg + bj]
  const int strips = P / 32;
  const int ccn = C / 2;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int PW = P / 2 + 1;
  if (threadIdx.x < 64) cacc[threadIdx.x] = 0.f;
  __syncthreads();
  for (int64_t item = blockIdx.x; item < items; item += gridDim.x) {
    const int strip = (int)(item % strips);
    const int64_t pair = item / strips;
    const int n = (int)(pair / ccn), cc = (int)(pair % ccn);
    const int t4 = strip * 32 + (lane >> 2) * 4;
    const int f4 = warp * 16 + (lane & 3) * 4;
    if (f4 < P) {
      float v[2][4][4];                      // [channel][row t4-1+r][column f4-1+c]
      float part[2] = {0.f, 0.f};
#pragma unroll
      for (int ch = 0; ch < 2; ++ch) {
        const int64_t plane = ((int64_t)n * C + 2 * cc + ch) * P * (int64_t)P;
        const float* ap = g1p + plane;
        const float* bp = gT + plane;
        const float* cp = gF + plane;
        float w[4][4];                       // w[c][r] = gF[f4-1+c][t4-1+r]
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int f = f4 - 1 + c;
          if (f >= 0) {
            const float* q = cp + (int64_t)f * P + t4;
            const float4 d = __ldg(reinterpret_cast<const float4*>(q));
            w[c][1] = d.x; w[c][2] = d.y; w[c][3] = d.z;
            w[c][0] = t4 > 0 ? __ldg(q - 1) : 0.f;
          } else {
            w[c][0] = w[c][1] = w[c][2] = w[c][3] = 0.f;
          }
        }
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const int t = t4 - 1 + r;
          if (t >= 0) {
            const int64_t off = (int64_t)t * P + f4;
            const float4 a = __ldg(reinterpret_cast<const float4*>(ap + off));
            const float4 b = __ldg(reinterpret_cast<const float4*>(bp + off));
            v[ch][r][1] = a.x - 0.5f * (b.x + w[1][r]); v[ch][r][2] = a.y - 0.5f * (b.y + w[2][r]);
            v[ch][r][3] = a.z - 0.5f * (b.z + w[3][r]);
            v[ch][r][0] = f4 > 0 ? __ldg(ap + off - 1) - 0.5f * (__ldg(bp + off - 1) + w[0][r]) : 0.f;
            part[ch] += (v[ch][r][0] + v[ch][r][1]) + (v[ch][r][2] + v[ch][r][3]);
          } else {
            v[ch][r][0] = v[ch][r][1] = v[ch][r][2] = v[ch][r][3] = 0.f;
          }
        }
      }
#pragma unroll
      for (int bi = 0; bi < 2; ++bi) {
#pragma unroll
        for (int bj = 0; bj < 2; ++bj) {
          float c[8];
#pragma unroll
          for (int ch = 0; ch < 2; ++ch) {
            c[ch * 4 + 0] = v[ch][2 * bi][2 * bj];     c[ch * 4 + 1] = v[ch][2 * bi][2 * bj + 1];
            c[ch * 4 + 2] = v[ch][2 * bi + 1][2 * bj]; c[ch * 4 + 3] = v[ch][2 * bi + 1][2 * bj + 1];
          }
          uint4 hh, ll;
          split8(c, hh, ll);
          stg[warp][0][2 * (lane >> 2) + bi][2 * (lane & 3) + bj] = hh;
          stg[warp][1][2 * (lane >> 2) + bi][2 * (lane & 3) + bj] = ll;
        }
      }
      __syncwarp();
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int idx = it * 32 + lane;
        const int half = idx >> 7, r = (idx >> 3) & 15, cx = idx & 7;
        const int64_t q = ((int64_t)n * PW + (strip * 16 + r)) * PW + (warp * 8 + cx);
        *reinterpret_cast<uint4*>(planes + (size_t)half * half_bytes + ((size_t)cc * (size_t)Qs + (size_t)q) * 16) = stg[warp][half][r][cx];
      }
      __syncwarp();
      // ---- bottom / right edge: pixel row P-1 / column P-1 pair with the zero beyond the map
      const bool bot = t4 == P - 4, rgt = f4 == P - 4;
      if (bot || rgt) {
        const int64_t pl0 = ((int64_t)n * C + 2 * cc) * P * (int64_t)P, pl1 = pl0 + (int64_t)P * P;
        if (bot) {
          // blocks (by = P/2, bx = f4/2 + bj): pixel row P-1, columns f4-1+2bj, f4+2bj
#pragma unroll
          for (int bj = 0; bj < 2; ++bj) {
            float c[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int f = f4 - 1 + 2 * bj + e;
              if (f >= 0) {
                c[e] = gx1_at(g1p + pl0, gT + pl0, gF + pl0, P, P - 1, f);
                c[4 + e] = gx1_at(g1p + pl1, gT + pl1, gF + pl1, P, P - 1, f);
                part[0] += c[e]; part[1] += c[4 + e];
              }
            }
            store_chunk(planes, half_bytes, Qs, cc, ((int64_t)n * PW + P / 2) * PW + (f4 / 2 + bj), c);
          }
        }
        if (rgt) {
          // blocks (by = t4/2 + bi, bx = P/2): pixel column P-1, rows t4-1+2bi, t4+2bi
#pragma unroll
          for (int bi = 0; bi < 2; ++bi) {
            float c[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int t = t4 - 1 + 2 * bi + e;
              if (t >= 0) {
                c[2 * e] = gx1_at(g1p + pl0, gT + pl0, gF + pl0, P, t, P - 1);
                c[4 + 2 * e] = gx1_at(g1p + pl1, gT + pl1, gF + pl1, P, t, P - 1);
                part[0] += c[2 * e]; part[1] += c[4 + 2 * e];
              }
            }
            store_chunk(planes, half_bytes, Qs, cc, ((int64_t)n * PW + (t4 / 2 + bi)) * PW + P / 2, c);
          }
        }
        if (bot && rgt) {
          float c[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
          c[0] = gx1_at(g1p + pl0, gT + pl0, gF + pl0, P, P - 1, P - 1);
          c[4] = gx1_at(g1p + pl1, gT + pl1, gF + pl1, P, P - 1, P - 1);
          part[0] += c[0]; part[1] += c[4];
          store_chunk(planes, half_bytes, Qs, cc, ((int64_t)n * PW + P / 2) * PW + P / 2, c);
        }
      }
      // per-channel sums: every lane of the warp reaches this point
      if (db1 != nullptr) {
#pragma unroll
        for (int ch = 0; ch < 2; ++ch) {
          const float p = warp_sum(part[ch]);
          if (lane == 0) atomicAdd(&cacc[2 * cc + ch], p);
        }
      }
    }
  }
  if (db1 != nullptr) {
    __syncthreads();
    if (threadIdx.x < C) atomicAdd(db1 + threadIdx.x, cacc[threadIdx.x]);
  }
}

// src/kharmonic_lofar.py:150-158 + its gradient (lshm_cascade_losses_upd), writing the gradients w.r.t. the two 1-D
// reconstructions as the pad-0 operand planes their nets' last transposed convs consume (window j = samples
// [4j, 4j+3] of the flattened map: a row of the thread's 4 x 4 block for the time-axis net, a column for the
// frequency-axis net).  Same micro-tile scheme; x3f is read along t (it is stored transposed), nothing goes through
// shared memory.  Persistent blocks: the seven loss sums and the bias-gradient sums are flushed once per block.
// UPD: apply the deferred multiplier update first; YZ: the multipliers are identically zero (a new minibatch,
// src/kharmonic_lofar.py:128-130) and are not read - with UPD the kernel then writes y_i = rho r_i.
template <bool UPD, bool YZ>
__global__ void __launch_bounds__(256, 2)
cascade_losses_planes_kernel(const float* __restrict__ x, const float* __restrict__ x1, const float* __restrict__ x2,
                             const float* __restrict__ x3f, float* __restrict__ y1, float* __restrict__ y2,
                             float* __restrict__ y3, float rho, float inv_n, int C, int P, int64_t items,
                             double* __restrict__ sums, float* __restrict__ g1p, uint8_t* __restrict__ p2,
                             uint8_t* __restrict__ p3, size_t half_bytes, int64_t Qs, float* __restrict__ db2,
                             float* __restrict__ db3) {
  __shared__ float cacc[2][64];
  __shared__ double red[8];
  const int strips = P / 32;
  const int ccn = C / 2;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x < 128) (&cacc[0][0])[threadIdx.x] = 0.f;
  __syncthreads();
  float s[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  const int64_t l = (int64_t)P * P / 4;
  for (int64_t item = blockIdx.x; item < items; item += gridDim.x) {
    const int strip = (int)(item % strips);
    const int64_t pair = item / strips;
    const int n = (int)(pair / ccn), cc = (int)(pair % ccn);
    const int t4 = strip * 32 + (lane >> 2) * 4;
    const int f4 = warp * 16 + (lane & 3) * 4;
    if (f4 >= P) continue;
    float g3v[2][4][4];                      // d/dx3 of the block, [channel][t][f]
    float xt[2][4][4];                       // x3 of the block read from the transposed tensor: [channel][f][t]
    float sb2[2] = {0.f, 0.f}, sb3[2] = {0.f, 0.f};
#pragma unroll
    for (int ch = 0; ch < 2; ++ch) {
      const float* q = x3f + ((int64_t)n * C + 2 * cc + ch) * P * (int64_t)P + (int64_t)f4 * P + t4;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float4 d = __ldg(reinterpret_cast<const float4*>(q + (int64_t)k * P));
        xt[ch][k][0] = d.x; xt[ch][k][1] = d.y; xt[ch][k][2] = d.z; xt[ch][k][3] = d.w;
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float g2row[8];
#pragma unroll
      for (int ch = 0; ch < 2; ++ch) {
        const int64_t off = (((int64_t)n * C + 2 * cc + ch) * P + (t4 + j)) * (int64_t)P + f4;
        const float4 X = __ldg(reinterpret_cast<const float4*>(x + off));
        const float4 A1 = __ldg(reinterpret_cast<const float4*>(x1 + off));
        const float4 A2 = __ldg(reinterpret_cast<const float4*>(x2 + off));
        float4 M1 = make_float4(0.f, 0.f, 0.f, 0.f), M2 = M1, M3 = M1;
        if (!YZ) {
          M1 = *reinterpret_cast<const float4*>(y1 + off);
          M2 = *reinterpret_cast<const float4*>(y2 + off);
          M3 = *reinterpret_cast<const float4*>(y3 + off);
        }
        const float xv[4] = {X.x, X.y, X.z, X.w}, a1[4] = {A1.x, A1.y, A1.z, A1.w}, a2[4] = {A2.x, A2.y, A2.z, A2.w};
        float m1[4] = {M1.x, M1.y, M1.z, M1.w}, m2[4] = {M2.x, M2.y, M2.z, M2.w}, m3[4] = {M3.x, M3.y, M3.z, M3.w};
        float g1[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float a3 = xt[ch][k][j];
          const float r0 = a1[k] + a2[k] + a3 - xv[k];
          const float r1 = xv[k] - a1[k];
          const float x11 = 0.5f * r1;
          const float r2 = x11 - a2[k], r3 = x11 - a3;
          if (UPD) { m1[k] = fmaf(rho, r1, m1[k]); m2[k] = fmaf(rho, r2, m2[k]); m3[k] = fmaf(rho, r3, m3[k]); }
          s[0] = fmaf(r0, r0, s[0]);
          s[1] = fmaf(m1[k], r1, s[1]); s[2] = fmaf(r1, r1, s[2]);
          s[3] = fmaf(m2[k], r2, s[3]); s[4] = fmaf(r2, r2, s[4]);
          s[5] = fmaf(m3[k], r3, s[5]); s[6] = fmaf(r3, r3, s[6]);
          const float e2 = m2[k] + rho * r2, e3 = m3[k] + rho * r3;
          const float v2 = (2.f * r0 - e2) * inv_n, v3 = (2.f * r0 - e3) * inv_n;
          g2row[ch * 4 + k] = v2;
          g3v[ch][j][k] = v3;
          g1[k] = (2.f * r0 - m1[k] - rho * r1 - 0.5f * (e2 + e3)) * inv_n;
          sb2[ch] += v2; sb3[ch] += v3;
        }
        if (UPD) {
          *reinterpret_cast<float4*>(y1 + off) = make_float4(m1[0], m1[1], m1[2], m1[3]);
          *reinterpret_cast<float4*>(y2 + off) = make_float4(m2[0], m2[1], m2[2], m2[3]);
          *reinterpret_cast<float4*>(y3 + off) = make_float4(m3[0], m3[1], m3[2], m3[3]);
        }
        *reinterpret_cast<float4*>(g1p + off) = make_float4(g1[0], g1[1], g1[2], g1[3]);
      }
      store_chunk(p2, half_bytes, Qs, cc, (int64_t)n * l + ((int64_t)(t4 + j) * P + f4) / 4, g2row);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float c[8] = {g3v[0][0][k], g3v[0][1][k], g3v[0][2][k], g3v[0][3][k],
                          g3v[1][0][k], g3v[1][1][k], g3v[1][2][k], g3v[1][3][k]};
      store_chunk(p3, half_bytes, Qs, cc, (int64_t)n * l + ((int64_t)(f4 + k) * P + t4) / 4, c);
    }
    if (db2 != nullptr) {
#pragma unroll
      for (int ch = 0; ch < 2; ++ch) {
        const float a = warp_sum(sb2[ch]), b = warp_sum(sb3[ch]);
        if (lane == 0) { atomicAdd(&cacc[0][2 * cc + ch], a); atomicAdd(&cacc[1][2 * cc + ch], b); }
      }
    }
  }
  __syncthreads();
  if (db2 != nullptr && threadIdx.x < C) {
    atomicAdd(db2 + threadIdx.x, cacc[0][threadIdx.x]); atomicAdd(db3 + threadIdx.x, cacc[1][threadIdx.x]);
  }
#pragma unroll
  for (int q = 0; q < 7; ++q) {
    const double v = warp_sum((double)s[q]);
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0.0;
      for (int w = 0; w < 8; ++w) t += red[w];
      atomicAdd(sums + q, t);
    }
  }
}

}  // namespace
}  // namespace lshm

using namespace lshm;

extern "C" {

int lshm_planes_bytes(int dim, int64_t N, int Bc, int h, int w_, int64_t* bytes) {
  LSHM_REQUIRE(bytes && (dim == 1 || dim == 2) && N >= 0 && Bc > 0 && h > 0 && w_ > 0, "lshm_planes_bytes: bad arguments");
  const PlaneGeom g = plane_geom(dim, N, Bc, h, w_);
  *bytes = (int64_t)(2 * g.half_bytes);
  return LSHM_OK;
}

int lshm_stage_planes2d(const float* big, int64_t big_ns, void* planes, int64_t N, int Bc, int h, int w_,
                        lshm_stream_t stream) {
  LSHM_REQUIRE(big && planes && N >= 0 && Bc > 0 && (Bc & 3) == 0 && h > 0 && w_ > 0, "lshm_stage_planes2d: bad arguments");
  LSHM_REQUIRE((reinterpret_cast<uintptr_t>(planes) & 15) == 0, "lshm_stage_planes2d: planes must be 16-byte aligned");
  if (N == 0) return LSHM_OK;
  const PlaneGeom g = plane_geom(2, N, Bc, h, w_);
  LSHM_REQUIRE(g.Q < (1LL << 31) - 4096, "lshm_stage_planes2d: too many positions for one call");
  const int64_t total = g.Q * g.chunks;
  const int64_t blocks = std::min<int64_t>(ceil_div(total, 256), (int64_t)sm_count() * 16);
  stage2d_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(big, big_ns, reinterpret_cast<uint8_t*>(planes), g.half_bytes,
      Bc, h, w_, g.Q, g.Qs, g.chunks, make_fastdiv((uint32_t)((h + 1) * (w_ + 1))), make_fastdiv((uint32_t)(w_ + 1)));
  LSHM_CHECK_LAUNCH("lshm_stage_planes2d");
  return LSHM_OK;
}

int lshm_stage_planes1d(const float* big, int64_t big_ns, void* planes, int64_t N, int Bc, int l, int pad,
                        lshm_stream_t stream) {
  LSHM_REQUIRE(big && planes && N >= 0 && Bc > 0 && (Bc & 3) == 0 && l > 0 && (pad == 0 || pad == 1), "lshm_stage_planes1d: bad arguments");
  LSHM_REQUIRE((reinterpret_cast<uintptr_t>(planes) & 15) == 0, "lshm_stage_planes1d: planes must be 16-byte aligned");
  if (N == 0) return LSHM_OK;
  const PlaneGeom g = plane_geom(1, N, Bc, 1, l);
  LSHM_REQUIRE(g.Q < (1LL << 31) - 4096, "lshm_stage_planes1d: too many positions for one call");
  const int64_t total = g.Q * g.chunks;
  const int64_t blocks = std::min<int64_t>(ceil_div(total, 256), (int64_t)sm_count() * 16);
  stage1d_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(big, big_ns, reinterpret_cast<uint8_t*>(planes), g.half_bytes,
      Bc, l, pad, g.Q, g.Qs, g.chunks, make_fastdiv((uint32_t)l));
  LSHM_CHECK_LAUNCH("lshm_stage_planes1d");
  return LSHM_OK;
}

int lshm_residual_split_planes(const float* x, const float* x1, void* planesT, void* planesF,
                               int64_t N, int C, int P, lshm_stream_t stream) {
  LSHM_REQUIRE(x && x1 && planesT && planesF, "lshm_residual_split_planes: null pointer");
  LSHM_REQUIRE(N >= 0 && C > 0 && (C & 3) == 0 && P > 0 && P % PT == 0, "lshm_residual_split_planes: need C%%4==0 and P%%32==0");
  if (N == 0) return LSHM_OK;
  const PlaneGeom g = plane_geom(1, N, C, 1, P * P / 4);
  LSHM_REQUIRE(P <= 128, "lshm_residual_split_planes: P > 128 is not supported");
  const int64_t blocks = N * (C / 2) * (int64_t)(P / 32);
  LSHM_REQUIRE(blocks < (1LL << 31), "lshm_residual_split_planes: batch too large for one call");
  residual_split_planes_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(
      x, x1, reinterpret_cast<uint8_t*>(planesT), reinterpret_cast<uint8_t*>(planesF), g.half_bytes, C, P, g.Qs);
  LSHM_CHECK_LAUNCH("lshm_residual_split_planes");
  return LSHM_OK;
}

int lshm_cascade_combine_planes(const float* g1p, const float* gT, const float* gF, void* planes,
                                int64_t N, int C, int P, float* db1, lshm_stream_t stream) {
  LSHM_REQUIRE(g1p && gT && gF && planes, "lshm_cascade_combine_planes: null pointer");
  LSHM_REQUIRE(N >= 0 && C > 0 && (C & 3) == 0 && C <= 64 && P > 0 && P % PT == 0, "lshm_cascade_combine_planes: need C%%4==0, C<=64 and P%%32==0");
  if (N == 0) return LSHM_OK;
  if (db1) LSHM_CUDA(cudaMemsetAsync(db1, 0, sizeof(float) * C, as_stream(stream)), "lshm_cascade_combine_planes");
  const PlaneGeom g = plane_geom(2, N, C, P / 2, P / 2);
  LSHM_REQUIRE(P <= 128, "lshm_cascade_combine_planes: P > 128 is not supported");
  const int64_t items = N * (C / 2) * (int64_t)(P / 32);
  const int64_t blocks = std::min<int64_t>(items, (int64_t)sm_count() * 8);
  combine_planes_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(g1p, gT, gF, reinterpret_cast<uint8_t*>(planes),
                                                                          g.half_bytes, C, P, g.Qs, items, db1);
  LSHM_CHECK_LAUNCH("lshm_cascade_combine_planes");
  return LSHM_OK;
}

int lshm_cascade_losses_planes(const float* x, const float* x1, const float* x2, const float* x3f,
                               float* y1, float* y2, float* y3, float rho, int update_y,
                               int64_t N, int C, int P, float grad_scale, double* sums,
                               float* g1p, void* planes2, void* planes3, float* db2, float* db3, lshm_stream_t stream) {
  LSHM_REQUIRE(x && x1 && x2 && x3f && y1 && y2 && y3 && sums && g1p && planes2 && planes3, "lshm_cascade_losses_planes: null pointer");
  LSHM_REQUIRE((db2 == nullptr) == (db3 == nullptr), "lshm_cascade_losses_planes: db2/db3 come together");
  LSHM_REQUIRE(N >= 0 && C > 0 && (C & 3) == 0 && C <= 64 && P > 0 && P % PT == 0 && P <= 128,
               "lshm_cascade_losses_planes: need C%%4==0, C<=64, P%%32==0, P<=128");
  if (N == 0) return LSHM_OK;
  cudaStream_t st = as_stream(stream);
  if (db2) {
    LSHM_CUDA(cudaMemsetAsync(db2, 0, sizeof(float) * C, st), "lshm_cascade_losses_planes");
    LSHM_CUDA(cudaMemsetAsync(db3, 0, sizeof(float) * C, st), "lshm_cascade_losses_planes");
  }
  const PlaneGeom g = plane_geom(1, N, C, 1, P * P / 4);
  const int64_t items = N * (C / 2) * (int64_t)(P / 32);
  const int64_t blocks = std::min<int64_t>(items, (int64_t)sm_count() * 2);
#define LSHM_CLP(U, Z) cascade_losses_planes_kernel<U, Z><<<(unsigned)blocks, 256, 0, st>>>(x, x1, x2, x3f, y1, y2, y3, rho, grad_scale, \
      C, P, items, sums, g1p, reinterpret_cast<uint8_t*>(planes2), reinterpret_cast<uint8_t*>(planes3), g.half_bytes, g.Qs, db2, db3)
  switch (update_y & 3) {
    case 0: LSHM_CLP(false, false); break;
    case 1: LSHM_CLP(true, false); break;
    case 2: LSHM_CLP(false, true); break;
    default: LSHM_CLP(true, true); break;
  }
#undef LSHM_CLP
  LSHM_CHECK_LAUNCH("lshm_cascade_losses_planes");
  return LSHM_OK;
}

}  // extern "C"
