// Bias gradients: db[c] = sum over samples and positions of g[n,c,:] (autograd of the bias add in
// every conv / transposed conv, /root/reference/src/lofar_models.py:73-78,:93-98).
//
// Rows (n,c) are contiguous and read as float4 with many loads in flight, no per-element index math.
#include <algorithm>
#include "common.cuh"

namespace lshm {
namespace {

__device__ __forceinline__ float warp_sum_active(float v) {   // all 32 lanes active
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// long rows: persistent warps walk (row, 8 KB chunk) items, 16 independent 16-byte loads per lane in
// flight, per-channel partial sums in shared memory, one global atomic per (block, channel).  The
// first version ran one short-lived block per row (4-64 KB): block start-up, the two-barrier block
// reduction and the atomic were a large part of every block's life (1.9 TB/s over the step's 36 calls).
constexpr int CS_CHUNK = 2048;                  // floats per work item
__global__ void __launch_bounds__(256)
channel_sum_rows_kernel(const float* __restrict__ g, int64_t g_ns, float* __restrict__ db, int64_t rows, int Cn,
                        int64_t len, int chunks_per_row, int vec_ok) {
  extern __shared__ float acc[];                // [Cn]
  for (int c = threadIdx.x; c < Cn; c += blockDim.x) acc[c] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int64_t items = rows * chunks_per_row;
  const int64_t nwarps = (int64_t)gridDim.x * 8;
  for (int64_t item = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); item < items; item += nwarps) {
    const int64_t row = item / chunks_per_row;
    const int ck = (int)(item - row * chunks_per_row);
    const int64_t n = row / Cn;
    const int c = (int)(row - n * Cn);
    const float* p = g + n * g_ns + (int64_t)c * len + (int64_t)ck * CS_CHUNK;
    const int cnt = (int)min((int64_t)CS_CHUNK, len - (int64_t)ck * CS_CHUNK);
    float s = 0.f;
    if (vec_ok) {
      const int n4 = cnt >> 2;
      float4 v[CS_CHUNK / 128];
#pragma unroll
      for (int u = 0; u < CS_CHUNK / 128; ++u) {
        const int i = lane + u * 32;
        v[u] = i < n4 ? ld_nc_f4(p + 4 * i) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < CS_CHUNK / 128; ++u) s += (v[u].x + v[u].y) + (v[u].z + v[u].w);
      for (int j = (n4 << 2) + lane; j < cnt; j += 32) s += __ldg(p + j);
    } else {
      for (int j = lane; j < cnt; j += 32) s += __ldg(p + j);
    }
    s = warp_sum(s);
    if (lane == 0) atomicAdd(&acc[c], s);
  }
  __syncthreads();
  for (int c = threadIdx.x; c < Cn; c += blockDim.x) atomicAdd(db + c, acc[c]);
}

// short rows (the deep layers: 4-256 values per row, 10^5 rows): every sample is one flat array (samples may
// be strided), a thread sums 16 bytes and adds them to its row's channel in shared memory; persistent blocks, one
// global atomic per (block, channel).  One warp per row in short-lived blocks (below, kept for strided or
// unaligned input) was 44 us per call in the ncu launch list: 25k blocks of 8 nearly idle warps.
__global__ void __launch_bounds__(256)
channel_sum_flat_kernel(const float* __restrict__ g, int64_t g_ns, float* __restrict__ db, int64_t n4, int Cn,
                        int len4) {
  extern __shared__ float acc[];                // [Cn]
  // (n4 is then a multiple of 32 too, so whole warps run every iteration)
  const bool warp_rows = (len4 & 31) == 0;
  for (int c = threadIdx.x; c < Cn; c += blockDim.x) acc[c] = 0.f;
  __syncthreads();
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t per_n = (int64_t)Cn * len4;     // 16-byte pieces per sample (samples may be strided: g_ns)
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const int64_t n = i / per_n;
    const int rem = (int)(i - n * per_n);
    const float4 v = ld_nc_f4(g + n * g_ns + 4 * (int64_t)rem);
    float sv = (v.x + v.y) + (v.z + v.w);
    if (warp_rows) {                            // rows of a multiple of 32 pieces: the warp's 32 pieces share a row
      sv = warp_sum_active(sv);
      if ((threadIdx.x & 31) == 0) atomicAdd(&acc[rem / len4], sv);
    } else {
      atomicAdd(&acc[rem / len4], sv);
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < Cn; c += blockDim.x) atomicAdd(db + c, acc[c]);
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

// dst += src (the gradients of a second micro-batch join the flat gradient buffer)
__global__ void vec_add_kernel(float* __restrict__ dst, const float* __restrict__ src, int64_t n4, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n4) {
    float4 a = reinterpret_cast<float4*>(dst)[i];
    const float4 b = reinterpret_cast<const float4*>(src)[i];
    a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    reinterpret_cast<float4*>(dst)[i] = a;
  }
  if (i == 0) for (int64_t j = n4 * 4; j < n; ++j) dst[j] += src[j];
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
    const int cpr = (int)ceil_div(len, (int64_t)CS_CHUNK);
    const int64_t items = rows * cpr;
    const int64_t want = ceil_div(items, (int64_t)8);
    const int grid = (int)std::min<int64_t>(want, (int64_t)sm_count() * 6);
    channel_sum_rows_kernel<<<grid, 256, sizeof(float) * Cn, st>>>(g, g_ns, db, rows, Cn, len, cpr, vec_ok);
  } else if ((reinterpret_cast<uintptr_t>(g) & 15) == 0 && (len & 3) == 0 && (g_ns & 3) == 0 && (int64_t)Cn * len < (1 << 30)) {
    const int64_t n4 = rows * (len >> 2);
    const int grid = (int)std::min<int64_t>(ceil_div(n4, (int64_t)256 * 4), (int64_t)sm_count() * 6);
    channel_sum_flat_kernel<<<grid, 256, sizeof(float) * Cn, st>>>(g, g_ns, db, n4, Cn, (int)(len >> 2));
  } else {
    channel_sum_short_kernel<<<(unsigned)ceil_div(rows, 8), 256, 0, st>>>(g, g_ns, db, rows, Cn, (int)len);
  }
  LSHM_CHECK_LAUNCH("lshm_channel_sum");
  return LSHM_OK;
}

int lshm_vec_add(float* dst, const float* src, int64_t n, lshm_stream_t stream) {
  LSHM_REQUIRE(dst && src && n >= 0, "lshm_vec_add: bad arguments");
  LSHM_REQUIRE(((reinterpret_cast<uintptr_t>(dst) | reinterpret_cast<uintptr_t>(src)) & 15) == 0, "lshm_vec_add: buffers must be 16-byte aligned");
  if (n == 0) return LSHM_OK;
  const int64_t n4 = n >> 2;
  vec_add_kernel<<<(unsigned)std::max<int64_t>(1, ceil_div(n4, 256)), 256, 0, as_stream(stream)>>>(dst, src, n4, n);
  LSHM_CHECK_LAUNCH("lshm_vec_add");
  return LSHM_OK;
}

}  // extern "C"
