// Weight-image management for the tensor-core conv kernels (see include/lshm.h).
#include "common.cuh"

namespace lshm {
size_t down_image_bytes(int dim, int A, int Bc);
int prep_down_image(const float* w, int dim, int A, int Bc, void* img, cudaStream_t st);
size_t up_image_bytes(int dim, int A, int Bc);
int prep_up_image(const float* w, int dim, int A, int Bc, void* img, cudaStream_t st);
}  // namespace lshm

using namespace lshm;

extern "C" {

int lshm_conv_image_bytes(int dim, int A, int Bc, int which, int64_t* bytes) {
  LSHM_REQUIRE(bytes && (dim == 1 || dim == 2) && A > 0 && Bc > 0 && (which == 0 || which == 1),
               "lshm_conv_image_bytes: bad arguments");
  *bytes = which == 0 ? (int64_t)down_image_bytes(dim, A, Bc) : (int64_t)up_image_bytes(dim, A, Bc);
  return LSHM_OK;
}

int lshm_conv_prep(const float* w, int dim, int A, int Bc, void* down_img, void* up_img, lshm_stream_t stream) {
  LSHM_REQUIRE(w && (dim == 1 || dim == 2) && A > 0 && Bc > 0, "lshm_conv_prep: bad arguments");
  if (down_img)
    if (int rc = prep_down_image(w, dim, A, Bc, down_img, as_stream(stream))) return rc;
  if (up_img)
    if (int rc = prep_up_image(w, dim, A, Bc, up_img, as_stream(stream))) return rc;
  return LSHM_OK;
}

}  // extern "C"
