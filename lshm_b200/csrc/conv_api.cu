// Weight-image management for the tensor-core conv kernels (see include/lshm.h).
#include "conv_geom.cuh"
#include "../../include/lshm_selftest.h"

namespace lshm {
namespace {

__global__ void prep_single_kernel(const float* __restrict__ w, int dim, int A, int Bc, int which, int64_t total,
                                   uint8_t* __restrict__ img) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  if (which == 0) prep_down_chunk(w, dim, A, Bc, down_geom(dim, A, Bc), idx, img);
  else prep_up_chunk(w, dim, A, Bc, up_geom(dim, A, Bc), idx, img);
}

// table record (8 x int64): w pointer, image pointer, dim, A, Bc, which, chunk count, unused
__global__ void prep_batch_kernel(const int64_t* __restrict__ table) {
  const int64_t* rec = table + 8 * (int64_t)blockIdx.y;
  const float* w = reinterpret_cast<const float*>(rec[0]);
  uint8_t* img = reinterpret_cast<uint8_t*>(rec[1]);
  const int dim = (int)rec[2], A = (int)rec[3], Bc = (int)rec[4], which = (int)rec[5];
  const int64_t total = rec[6];
  if (which == 0) {
    const DownGeom g = down_geom(dim, A, Bc);
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x)
      prep_down_chunk(w, dim, A, Bc, g, idx, img);
  } else {
    const UpGeom g = up_geom(dim, A, Bc);
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x)
      prep_up_chunk(w, dim, A, Bc, g, idx, img);
  }
}

}  // namespace
}  // namespace lshm

using namespace lshm;

extern "C" {

int lshm_conv_image_bytes(int dim, int A, int Bc, int which, int64_t* bytes) {
  LSHM_REQUIRE(bytes && (dim == 1 || dim == 2) && A > 0 && Bc > 0 && (which == 0 || which == 1),
               "lshm_conv_image_bytes: bad arguments");
  if (which == 0) { const DownGeom g = down_geom(dim, A, Bc); *bytes = (int64_t)(g.img * g.ntiles * g.KB); }
  else { const UpGeom g = up_geom(dim, A, Bc); *bytes = (int64_t)(g.img * g.ntiles * g.KB); }
  return LSHM_OK;
}

int lshm_conv_prep(const float* w, int dim, int A, int Bc, void* down_img, void* up_img, lshm_stream_t stream) {
  LSHM_REQUIRE(w && (dim == 1 || dim == 2) && A > 0 && Bc > 0, "lshm_conv_prep: bad arguments");
  cudaStream_t st = as_stream(stream);
  if (down_img) {
    const int64_t total = down_chunks(down_geom(dim, A, Bc));
    prep_single_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, st>>>(w, dim, A, Bc, 0, total, reinterpret_cast<uint8_t*>(down_img));
  }
  if (up_img) {
    const int64_t total = up_chunks(up_geom(dim, A, Bc));
    prep_single_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, st>>>(w, dim, A, Bc, 1, total, reinterpret_cast<uint8_t*>(up_img));
  }
  LSHM_CHECK_LAUNCH("lshm_conv_prep");
  return LSHM_OK;
}

int lshm_conv_prep_record(const float* w, int dim, int A, int Bc, int which, void* img, int64_t* record) {
  LSHM_REQUIRE(w && img && record && (dim == 1 || dim == 2) && A > 0 && Bc > 0 && (which == 0 || which == 1),
               "lshm_conv_prep_record: bad arguments");
  record[0] = (int64_t)reinterpret_cast<uintptr_t>(w);
  record[1] = (int64_t)reinterpret_cast<uintptr_t>(img);
  record[2] = dim; record[3] = A; record[4] = Bc; record[5] = which;
  record[6] = which == 0 ? down_chunks(down_geom(dim, A, Bc)) : up_chunks(up_geom(dim, A, Bc));
  record[7] = 0;
  return LSHM_OK;
}

int lshm_conv_prep_batch(const int64_t* table, int n, lshm_stream_t stream) {
  LSHM_REQUIRE(table && n >= 0, "lshm_conv_prep_batch: bad arguments");
  if (n == 0) return LSHM_OK;
  dim3 grid(48, (unsigned)n);
  prep_batch_kernel<<<grid, 256, 0, as_stream(stream)>>>(table);
  LSHM_CHECK_LAUNCH("lshm_conv_prep_batch");
  return LSHM_OK;
}

int lshm_fastdiv_check(int64_t d, const int64_t* n, int count, int64_t* mismatches) {
  LSHM_REQUIRE(d >= 1 && d < (1LL << 31) && n != nullptr && mismatches != nullptr && count >= 0,
               "lshm_fastdiv_check: bad arguments");
  const FastDiv f = make_fastdiv((uint32_t)d);
  int bad = 0;
  for (int i = 0; i < count; ++i) {
    const uint32_t v = (uint32_t)n[i];
    // host copy of fdiv(): __umulhi(m, v) = high word of the 64-bit product
    const uint32_t q = (uint32_t)((((uint64_t)f.m * v) >> 32) + v) >> f.l;
    if (n[i] < 0 || n[i] >= (1LL << 31) || q != v / (uint32_t)d) ++bad;
  }
  *mismatches = bad;
  return LSHM_OK;
}

}  // extern "C"
