// Bias gradients: db[c] = sum over samples and positions of g[n,c,:] (autograd of the bias add in
// every conv / transposed conv, /root/reference/src/lofar_models.py:73-78,:93-98).
#include "common.cuh"

namespace lshm {
namespace {

// ------------------------------------------------------------------ bias gradient --------
__global__ void __launch_bounds__(256)
channel_sum_kernel(const float* __restrict__ g, int64_t g_ns, float* __restrict__ db, int64_t N,
                   int Cn, int64_t len, int64_t chunk) {
  __shared__ float red[32];
  const int c = blockIdx.y;
  const int64_t total = N * len;
  const int64_t start = (int64_t)blockIdx.x * chunk, stop = min(start + chunk, total);
  float s = 0.f;
  for (int64_t idx = start + threadIdx.x; idx < stop; idx += blockDim.x) {
    const int64_t n = idx / len, r = idx - n * len;
    s += __ldg(g + n * g_ns + (int64_t)c * len + r);
  }
  s = block_sum<float>(s, red);
  if (threadIdx.x == 0) atomicAdd(db + c, s);
}

int64_t pick_chunk(int64_t total, int64_t combos, int threads) {
  // aim for ~16 blocks per SM overall, at least 4 and at most 64 pixels per thread
  const int64_t target_blocks = std::max<int64_t>(1, (int64_t)sm_count() * 16 / std::max<int64_t>(1, combos));
  int64_t ppt = ceil_div(total, target_blocks * threads);
  ppt = std::max<int64_t>(4, std::min<int64_t>(64, ppt));
  return ppt * threads;
}

}  // namespace
}  // namespace lshm

using namespace lshm;

extern "C" {

int lshm_channel_sum(const float* g, int64_t g_ns, float* db, int64_t N, int Cn, int64_t len,
                     lshm_stream_t stream) {
  LSHM_REQUIRE(g && db && N >= 0 && Cn > 0 && len > 0, "lshm_channel_sum: bad arguments");
  cudaStream_t st = as_stream(stream);
  LSHM_CUDA(cudaMemsetAsync(db, 0, sizeof(float) * Cn, st), "lshm_channel_sum");
  if (N == 0) return LSHM_OK;
  const int64_t total = N * len;
  const int64_t chunk = pick_chunk(total, Cn, 256);
  dim3 grid((unsigned)ceil_div(total, chunk), (unsigned)Cn);
  channel_sum_kernel<<<grid, 256, 0, st>>>(g, g_ns, db, N, Cn, len, chunk);
  LSHM_CHECK_LAUNCH("lshm_channel_sum");
  return LSHM_OK;
}

}  // extern "C"
