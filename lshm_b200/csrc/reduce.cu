// Bias gradients: db[c] = sum over samples and positions of g[n,c,:] (autograd of the bias add in
// every conv / transposed conv, /root/reference/src/lofar_models.py:73-78,:93-98).
//
// One block per group of rows (n,c): rows are contiguous, read as float4 with many loads in flight,
// no per-element index math; one atomic per (block, channel).
#include "common.cuh"

namespace lshm {
namespace {

// long rows: one block per (n, c) row segment
__global__ void __launch_bounds__(256)
channel_sum_rows_kernel(const float* __restrict__ g, int64_t g_ns, float* __restrict__ db, int Cn, int64_t len,
                        int vec_ok) {
  __shared__ float red[32];
  const int64_t row = blockIdx.x;             // n * Cn + c
  const int64_t n = row / Cn;
  const int c = (int)(row - n * Cn);
  const float* p = g + n * g_ns + (int64_t)c * len;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  if (vec_ok) {
    const float4* p4 = reinterpret_cast<const float4*>(p);
    const int64_t n4 = len >> 2;
    int64_t i = threadIdx.x;
    for (; i + 3 * 256 < n4; i += 4 * 256) {
      const float4 a = ld_nc_f4(reinterpret_cast<const float*>(p4 + i));
      const float4 b = ld_nc_f4(reinterpret_cast<const float*>(p4 + i + 256));
      const float4 cc = ld_nc_f4(reinterpret_cast<const float*>(p4 + i + 512));
      const float4 d = ld_nc_f4(reinterpret_cast<const float*>(p4 + i + 768));
      s0 += (a.x + a.y) + (a.z + a.w); s1 += (b.x + b.y) + (b.z + b.w);
      s2 += (cc.x + cc.y) + (cc.z + cc.w); s3 += (d.x + d.y) + (d.z + d.w);
    }
    for (; i < n4; i += 256) {
      const float4 a = ld_nc_f4(reinterpret_cast<const float*>(p4 + i));
      s0 += (a.x + a.y) + (a.z + a.w);
    }
    for (int64_t j = (n4 << 2) + threadIdx.x; j < len; j += 256) s1 += p[j];
  } else {
    for (int64_t j = threadIdx.x; j < len; j += 256) s0 += p[j];
  }
  const float s = block_sum<float>((s0 + s1) + (s2 + s3), red);
  if (threadIdx.x == 0) atomicAdd(db + c, s);
}

// short rows: one warp per (n, c) row, 8 rows per block
__global__ void __launch_bounds__(256)
channel_sum_short_kernel(const float* __restrict__ g, int64_t g_ns, float* __restrict__ db, int64_t rows, int Cn,
                         int len) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int64_t n = row / Cn;
  const int c = (int)(row - n * Cn);
  const float* p = g + n * g_ns + (int64_t)c * len;
  float s = 0.f;
  for (int j = lane; j < len; j += 32) s += __ldg(p + j);
  s = warp_sum(s);
  if (lane == 0) atomicAdd(db + c, s);
}

}  // namespace
}  // namespace lshm

using namespace lshm;

extern "C" {

int lshm_channel_sum(const float* g, int64_t g_ns, float* db, int64_t N, int Cn, int64_t len,
                     lshm_stream_t stream) {
  LSHM_REQUIRE(g && db && N >= 0 && Cn > 0 && len > 0, "lshm_channel_sum: bad arguments");
  LSHM_REQUIRE(N * Cn < (1LL << 31), "lshm_channel_sum: too many rows for one call");
  cudaStream_t st = as_stream(stream);
  LSHM_CUDA(cudaMemsetAsync(db, 0, sizeof(float) * Cn, st), "lshm_channel_sum");
  if (N == 0) return LSHM_OK;
  const int64_t rows = N * Cn;
  if (len >= 512) {
    const int vec_ok = ((reinterpret_cast<uintptr_t>(g) & 15) == 0 && (g_ns & 3) == 0 && (len & 3) == 0) ? 1 : 0;
    channel_sum_rows_kernel<<<(unsigned)rows, 256, 0, st>>>(g, g_ns, db, Cn, len, vec_ok);
  } else {
    channel_sum_short_kernel<<<(unsigned)ceil_div(rows, 8), 256, 0, st>>>(g, g_ns, db, rows, Cn, (int)len);
  }
  LSHM_CHECK_LAUNCH("lshm_channel_sum");
  return LSHM_OK;
}

}  // extern "C"
