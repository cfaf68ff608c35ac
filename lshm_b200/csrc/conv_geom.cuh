// Weight-image geometry and the per-chunk preparation code shared by the tensor-core conv kernels.
#pragma once
#include "tc_common.cuh"

namespace lshm {

// Division of a 31-bit index by a launch-time constant: q = (mulhi(m, n) + n) >> l with
// l = ceil(log2 d), m = floor(2^32 (2^l - d) / d) + 1 (Granlund-Montgomery; n < 2^31 keeps the sum in
// 32 bits).  The compiler's generic 32-bit division is ~20 instructions and the position -> (image,
// row, column) split is on every producer's and epilogue's critical path.
struct FastDiv { uint32_t d, m, l; };
inline FastDiv make_fastdiv(uint32_t d) {
  FastDiv f; f.d = d; f.l = 0;
  while ((1ull << f.l) < d) ++f.l;
  f.m = (uint32_t)((((1ull << f.l) - d) << 32) / d + 1);
  return f;
}
__device__ __forceinline__ uint32_t fdiv(uint32_t n, const FastDiv& f) { return (__umulhi(f.m, n) + n) >> f.l; }

struct DownGeom { int NT, KC, ntiles, KB, T; size_t img; };
struct UpGeom { int NT, KC, ntiles, KB, combos, ncols; size_t img; };

__host__ __device__ inline DownGeom down_geom(int dim, int A, int Bc) {
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

__host__ __device__ inline UpGeom up_geom(int dim, int A, int Bc) {
  UpGeom g;
  g.ncols = dim == 2 ? Bc : 4 * Bc;                 // GEMM N extent
  const int n16 = (g.ncols + 15) / 16 * 16;
  if (dim == 2) g.NT = n16 <= 16 ? 16 : (n16 <= 32 ? 32 : 48);
  else g.NT = n16 <= 16 ? 16 : (n16 <= 32 ? 32 : (n16 <= 48 ? 48 : 96));
  g.KC = (dim == 2 || A <= 16) ? 16 : 32;
  g.ntiles = (g.ncols + g.NT - 1) / g.NT;
  g.KB = ((A + 15) / 16 * 16 + g.KC - 1) / g.KC;
  g.combos = dim == 2 ? 16 : 1;
  g.img = (size_t)2 * g.combos * (g.KC / 8) * g.NT * 16;
  return g;
}

// "down" image: [ntile][kblock][half hi/lo][tap][chunk][a_local][8 x bf16]; idx = one 8-element chunk
__device__ __forceinline__ void prep_down_chunk(const float* __restrict__ w, int dim, int A, int Bc, const DownGeom& g,
                                                int64_t idx, uint8_t* __restrict__ img) {
  const int NT = g.NT, KC = g.KC, KB = g.KB, T = g.T, CC = KC / 8;
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
  const size_t blk = g.img;
  uint8_t* base = img + ((size_t)nt * KB + kb) * blk + (((size_t)tap * CC + cc) * NT + al) * 16;
  *reinterpret_cast<uint4*>(base) = hi;
  *reinterpret_cast<uint4*>(base + blk / 2) = lo;
}

// "up" image: [ntile][kblock][half][class*4+tap][chunk][n_local][8 x bf16 over a]
__device__ __forceinline__ void prep_up_chunk(const float* __restrict__ w, int dim, int A, int Bc, const UpGeom& g,
                                              int64_t idx, uint8_t* __restrict__ img) {
  const int NT = g.NT, KC = g.KC, KB = g.KB, combos = g.combos, CC = KC / 8;
  int64_t r = idx;
  const int nl = (int)(r % NT); r /= NT;
  const int cc = (int)(r % CC); r /= CC;
  const int combo = (int)(r % combos); r /= combos;
  const int kb = (int)(r % KB); r /= KB;
  const int nt = (int)r;
  const int col = nt * NT + nl;
  float v[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int ch = kb * KC + cc * 8 + e;
    float x = 0.f;
    if (ch < A) {
      if (dim == 2) {
        if (col < Bc) {
          const int cls = combo >> 2, tap = combo & 3;
          const int ry = cls >> 1, rx = cls & 1, d = tap >> 1, ee = tap & 1;
          const int ky = ry == 0 ? (d == 0 ? 1 : 3) : (d == 0 ? 0 : 2);
          const int kx = rx == 0 ? (ee == 0 ? 1 : 3) : (ee == 0 ? 0 : 2);
          x = w[(((int64_t)ch * Bc + col) * 4 + ky) * 4 + kx];
        }
      } else if (col < 4 * Bc) {
        x = w[(int64_t)ch * Bc * 4 + col];          // col = b*4 + t
      }
    }
    v[e] = x;
  }
  uint4 hi, lo;
  tc::split8(v, hi, lo);
  const size_t blk = g.img;
  uint8_t* base = img + ((size_t)nt * KB + kb) * blk + (((size_t)combo * CC + cc) * NT + nl) * 16;
  *reinterpret_cast<uint4*>(base) = hi;
  *reinterpret_cast<uint4*>(base + blk / 2) = lo;
}

__host__ __device__ inline int64_t down_chunks(const DownGeom& g) { return (int64_t)g.ntiles * g.KB * g.T * (g.KC / 8) * g.NT; }
__host__ __device__ inline int64_t up_chunks(const UpGeom& g) { return (int64_t)g.ntiles * g.KB * g.combos * (g.KC / 8) * g.NT; }

}  // namespace lshm
