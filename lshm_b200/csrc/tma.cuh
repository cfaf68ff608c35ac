// Tensor-TMA helpers for the "operand planes" of the conv kernels (sm_100a).
//
// Operand planes = an input-sized (level-0) tensor stored the way the tensor-core conv kernels consume it:
// the space-to-depth operand Z[q, c] of igemm_down.cu / igemm_wgrad.cu, already split into bf16 hi / lo
// halves, as two arrays  plane[half][chunk cc = c/8][position q][8 x bf16]  (same bytes as the fp32 tensor,
// + the zero halo row / column of the 2-D block grid).  A CTA's shared-memory tile
// [chunk][slot][8 x bf16] is then ONE box of a 3-D tensor map {8, Q, chunks}: the producer warps (gather,
// fp32 -> bf16 hi/lo split, shared-memory stores: 28 instructions per 8 values, the bound of the first-layer
// kernels in round 1) are replaced by one cp.async.bulk.tensor per half and stage; positions past the end
// of the tensor are zero-filled by the copy engine.
#pragma once
#include <cuda.h>
#include "tc_common.cuh"

namespace lshm {

// Q = positions, Qs = chunk stride in positions (Q rounded up to the 32-position rows the copy engine fetches; the
// <= 31 padding positions per chunk are never written: the buffer must be zero-filled once when it is allocated)
struct PlaneGeom { int64_t Q; int64_t Qs; int chunks; size_t half_bytes; };
constexpr int PLANE_ROW = 32;     // positions per tensor-map row (32 x 16 B = 512 B)

// dim 2: big map [N,Bc,2h,2w], Q = N (h+1)(w+1) block positions incl. the halo; dim 1: [N,Bc,4l], Q = N l.
__host__ __device__ inline PlaneGeom plane_geom(int dim, int64_t N, int Bc, int h, int w) {
  PlaneGeom g;
  g.Q = dim == 2 ? N * (int64_t)(h + 1) * (w + 1) : N * (int64_t)w;
  g.Qs = (g.Q + PLANE_ROW - 1) / PLANE_ROW * PLANE_ROW;
  g.chunks = (4 * Bc + 7) / 8;
  g.half_bytes = (size_t)g.chunks * (size_t)g.Qs * 16;
  return g;
}

// Host: tensor map over one half (hi or lo) of a plane buffer.  The half is described as
// {64 x 8-byte elements (= 32 positions, 512 B contiguous), Qs/32 rows, chunks}; box = {64, slots/32, box_chunks}, so a
// tile arrives as slots/32 * box_chunks rows of 512 B.  (A first version used {8 x bf16, Q, chunks} with 16-byte rows:
// the copy engine then issues one request per 16 B and sustained only ~0.7 rows per cycle and SM - the kernels were
// bound by the TMA request rate, profiles/r2_ncu_planes.md.)  `slots` must be a multiple of 32 and tiles start at
// multiples of 32 positions.  cuTensorMapEncodeTiled is fetched through the runtime (no link-time libcuda dependency).
int make_plane_tmap(CUtensorMap* m, const void* half_base, int64_t Qs, int chunks, int slots, int box_chunks);

namespace tc {

__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" :: "l"(reinterpret_cast<uint64_t>(m)) : "memory");
}

// 3-D tiled load global -> shared, completion (bytes) counted on `bar`
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* m, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
               :: "r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}

}  // namespace tc
}  // namespace lshm
