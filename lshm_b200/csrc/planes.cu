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
constexpr int PT = 32;

// x11 = (x - x1) / 2 written as the 1-D pad-1 planes of the time-axis net (flattened s = t*P + f) and of the
// frequency-axis net (flattened s = f*P + t): window j of a flattened map covers samples [4j-1, 4j+2].
// Block: (tile, channel pair); 256 threads.  The tile is extended by one leading column (time net) / one
// leading row (frequency net) so that every window that STARTS in the tile is complete.
__global__ void __launch_bounds__(256)
residual_split_planes_kernel(const float* __restrict__ x, const float* __restrict__ x1, uint8_t* __restrict__ pT,
                             uint8_t* __restrict__ pF, size_t half_bytes, int C, int P, int64_t Q, int chunks) {
  __shared__ float s[2][PT + 1][PT + 2];     // [channel][row t (+1 halo row above)][col f (+1 halo col left)]
  const int tpr = P / PT;
  const int tile = blockIdx.x % (tpr * tpr), pair = blockIdx.x / (tpr * tpr);
  const int ccn = C / 2;
  const int n = pair / ccn, cc = pair % ccn;
  const int t0 = (tile / tpr) * PT, f0 = (tile % tpr) * PT;
  const int tid = threadIdx.x;
  // load (PT+1) x (PT+1) values per channel: rows t0-1 .. t0+PT-1, cols f0-1 .. f0+PT-1.  The flattened
  // predecessor of (t, f=0) is (t-1, P-1) for the time net and of (f, t=0) is (f-1, P-1) for the frequency net:
  // those wrap-around values are fetched separately below, the halo here serves the in-row windows.
  for (int i = tid; i < 2 * (PT + 1) * (PT + 1); i += 256) {
    const int ch = i / ((PT + 1) * (PT + 1)), r = (i / (PT + 1)) % (PT + 1), c = i % (PT + 1);
    const int t = t0 - 1 + r, f = f0 - 1 + c;
    float v = 0.f;
    if (t >= 0 && f >= 0) {
      const int64_t off = (((int64_t)n * C + 2 * cc + ch) * P + t) * P + f;
      v = 0.5f * (__ldg(x + off) - __ldg(x1 + off));
    }
    s[ch][r][c] = v;
  }
  __syncthreads();
  const int64_t plane0 = ((int64_t)n * C + 2 * cc) * P * (int64_t)P;
  const int wpr = PT / 4;                    // windows per tile row
  // ---- time net: windows along f.  Window starting at sample s0 = t*P + 4m - 1 (m = window index in the row).
  for (int i = tid; i < PT * wpr; i += 256) {
    const int r = i / wpr, m = i % wpr;
    const int t = t0 + r, fw = f0 + 4 * m;   // window covers f = fw-1 .. fw+2
    float v[8];
#pragma unroll
    for (int ch = 0; ch < 2; ++ch) {
      float first = s[ch][r + 1][4 * m];     // (t, fw-1); for fw == 0 the predecessor is (t-1, P-1)
      if (fw == 0) {
        first = 0.f;
        if (t > 0) {
          const int64_t off = plane0 + (int64_t)ch * P * P + (int64_t)(t - 1) * P + (P - 1);
          first = 0.5f * (__ldg(x + off) - __ldg(x1 + off));
        }
      }
      v[ch * 4 + 0] = first;
      v[ch * 4 + 1] = s[ch][r + 1][4 * m + 1];
      v[ch * 4 + 2] = s[ch][r + 1][4 * m + 2];
      v[ch * 4 + 3] = s[ch][r + 1][4 * m + 3];
    }
    const int64_t q = (int64_t)n * (P * P / 4) + ((int64_t)t * P + fw) / 4;
    store_chunk(pT, half_bytes, Q, cc, q, v);
  }
  // ---- frequency net: flattened s = f*P + t, windows along t at fixed f.
  for (int i = tid; i < PT * wpr; i += 256) {
    const int c = i / wpr, m = i % wpr;      // consecutive threads -> consecutive windows of one f (contiguous stores)
    const int f = f0 + c, tw = t0 + 4 * m;   // window covers t = tw-1 .. tw+2
    float v[8];
#pragma unroll
    for (int ch = 0; ch < 2; ++ch) {
      float first = s[ch][4 * m][c + 1];     // (tw-1, f); for tw == 0 the predecessor is (P-1, f-1)
      if (tw == 0) {
        first = 0.f;
        if (f > 0) {
          const int64_t off = plane0 + (int64_t)ch * P * P + (int64_t)(P - 1) * P + (f - 1);
          first = 0.5f * (__ldg(x + off) - __ldg(x1 + off));
        }
      }
      v[ch * 4 + 0] = first;
      v[ch * 4 + 1] = s[ch][4 * m + 1][c + 1];
      v[ch * 4 + 2] = s[ch][4 * m + 2][c + 1];
      v[ch * 4 + 3] = s[ch][4 * m + 3][c + 1];
    }
    const int64_t q = (int64_t)n * (P * P / 4) + ((int64_t)f * P + tw) / 4;
    store_chunk(pF, half_bytes, Q, cc, q, v);
  }
}

// gx1 = g1p - 0.5 (gT + transpose(gF)) written as the 2-D planes of the 2-D net's last transposed conv
// (block (by,bx) = pixel rows 2by-1, 2by x columns 2bx-1, 2bx), plus the per-channel sums (bias gradient).
// Tiles are shifted by (-1,-1) so that they hold whole 2x2 blocks: tile (i,j) covers pixel rows 32i-1 .. 32i+30;
// a fifth tile row / column holds the last halo blocks (pixel row / column P-1 only).
__global__ void __launch_bounds__(256)
combine_planes_kernel(const float* __restrict__ g1p, const float* __restrict__ gT, const float* __restrict__ gF,
                      uint8_t* __restrict__ planes, size_t half_bytes, int C, int P, int64_t Q, int64_t items,
                      float* __restrict__ db1) {
  __shared__ float s[2][PT][PT + 1];
  __shared__ float tr[PT][PT + 1];
  __shared__ float cacc[64];                 // per-channel partial sums of this (persistent) block
  const int tpr = P / PT + 1;                // tiles per row incl. the halo tile
  const int ccn = C / 2;
  const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
  const int PW = P / 2 + 1;
  if (tid < 64) cacc[tid] = 0.f;
  for (int64_t item = blockIdx.x; item < items; item += gridDim.x) {
    const int tile = (int)(item % (tpr * tpr));
    const int64_t pair = item / (tpr * tpr);
    const int n = (int)(pair / ccn), cc = (int)(pair % ccn);
    const int ti = tile / tpr, tj = tile % tpr;
    const int r0 = ti * PT - 1, c0 = tj * PT - 1;      // first pixel row / column of the tile
#pragma unroll
    for (int ch = 0; ch < 2; ++ch) {
      const int64_t plane = ((int64_t)n * C + 2 * cc + ch) * P * (int64_t)P;
      __syncthreads();
      // transposed input: tr[a][b] = gF[f = c0 + a][t = r0 + b]
#pragma unroll
      for (int i = 0; i < PT; i += 8) {
        const int f = c0 + ty + i, t = r0 + tx;
        tr[ty + i][tx] = (f >= 0 && f < P && t >= 0 && t < P) ? __ldg(gF + plane + (int64_t)f * P + t) : 0.f;
      }
      __syncthreads();
      float part = 0.f;
#pragma unroll
      for (int i = 0; i < PT; i += 8) {
        const int t = r0 + ty + i, f = c0 + tx;
        float v = 0.f;
        if (t >= 0 && t < P && f >= 0 && f < P) {
          const int64_t off = plane + (int64_t)t * P + f;
          v = __ldg(g1p + off) - 0.5f * (__ldg(gT + off) + tr[tx][ty + i]);
        }
        s[ch][ty + i][tx] = v;
        part += v;
      }
      if (db1 != nullptr) {
        part = warp_sum(part);
        if (tx == 0) atomicAdd(&cacc[2 * cc + ch], part);
      }
    }
    __syncthreads();
    // 16 x 16 blocks per tile, one per thread
    const int byl = tid >> 4, bxl = tid & 15;
    const int by = ti * (PT / 2) + byl, bx = tj * (PT / 2) + bxl;
    if (by < PW && bx < PW) {
      float v[8];
#pragma unroll
      for (int ch = 0; ch < 2; ++ch) {
        v[ch * 4 + 0] = s[ch][2 * byl][2 * bxl];
        v[ch * 4 + 1] = s[ch][2 * byl][2 * bxl + 1];
        v[ch * 4 + 2] = s[ch][2 * byl + 1][2 * bxl];
        v[ch * 4 + 3] = s[ch][2 * byl + 1][2 * bxl + 1];
      }
      const int64_t q = ((int64_t)n * PW + by) * PW + bx;
      store_chunk(planes, half_bytes, Q, cc, q, v);
    }
  }
  if (db1 != nullptr) {
    __syncthreads();
    if (tid < C) atomicAdd(db1 + tid, cacc[tid]);
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
  const int64_t blocks = N * (C / 2) * (int64_t)(P / PT) * (P / PT);
  LSHM_REQUIRE(blocks < (1LL << 31), "lshm_residual_split_planes: batch too large for one call");
  residual_split_planes_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(
      x, x1, reinterpret_cast<uint8_t*>(planesT), reinterpret_cast<uint8_t*>(planesF), g.half_bytes, C, P, g.Qs, g.chunks);
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
  const int tpr = P / PT + 1;
  const int64_t items = N * (C / 2) * (int64_t)tpr * tpr;
  const int64_t blocks = std::min<int64_t>(items, (int64_t)sm_count() * 8);
  combine_planes_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(g1p, gT, gF, reinterpret_cast<uint8_t*>(planes),
                                                                          g.half_bytes, C, P, g.Qs, items, db1);
  LSHM_CHECK_LAUNCH("lshm_cascade_combine_planes");
  return LSHM_OK;
}

}  // extern "C"
