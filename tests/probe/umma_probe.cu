// tcgen05 probe (development aid, not part of the product library): validates the shared-memory
// descriptor conventions used by lshm_b200/csrc/igemm*.cu on real sm_100a hardware:
//   mode 0: A K-major, rows at a uniform 16-byte pitch (SBO=128), chunk stride LBO; optional
//           start-address shift by `shift` rows (the "shifted tap" trick);
//   mode 1: as 0 but LBO/SBO roles swapped in the descriptor (must FAIL if 0 is right);
//   mode 2: A MN-major: tile X[pos][c] (pos rows at 16 B pitch, 8 c per 16 B), M=c, K=pos;
//   mode 3: as 2 with LBO/SBO swapped.
// B is always K-major [N][K].  bf16 hi/lo split, 3 MMAs per K step, fp32 accumulate in TMEM.
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) |
         (1ull << 46);
}

__device__ __forceinline__ void mma_bf16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n"
               :: "r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}

struct Params { int mode, shift, N, K; };

__global__ void __launch_bounds__(128) probe_kernel(const float* A, const float* B, float* out, Params p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t tmem_base;
  __shared__ __align__(8) uint64_t bar;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int SLOTS = 128 + 96;                      // row slots per chunk column
  const int Npad = p.N;
  const int nchunk = p.K / 8;
  const int PSLOTS = p.K + 96;
  const int a_bytes = max(nchunk * SLOTS * 16, 16 * PSLOTS * 16);
  uint8_t* a_hi = smem;
  uint8_t* a_lo = a_hi + a_bytes;
  uint8_t* b_hi = a_lo + a_bytes;
  uint8_t* b_lo = b_hi + nchunk * Npad * 16;
  // MN-major mode: A tile indexed [pos][c]: pos = K index (p.K positions + shift slack), c = M (128)
  // stored as chunk column per 8 c: address = (c/8)*CH + pos*16 + (c%8)*2, CH = PSLOTS*16
  uint8_t* x_hi = a_hi;  // reuse: needs 16 chunk-columns * PSLOTS*16 <= nchunk*SLOTS*16 ? sized by host
  uint8_t* x_lo = a_lo;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base)), "r"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(&bar)), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  // ---- fill operand tiles (hi/lo split)
  if (p.mode < 2) {
    for (int idx = tid; idx < 128 * p.K; idx += 128) {
      const int r = idx / p.K, k = idx % p.K;
      const float v = A[r * p.K + k];
      const __nv_bfloat16 h = __float2bfloat16_rn(v);
      const __nv_bfloat16 l = __float2bfloat16_rn(v - __bfloat162float(h));
      const int off = (k / 8) * SLOTS * 16 + (r + p.shift) * 16 + (k % 8) * 2;
      *reinterpret_cast<__nv_bfloat16*>(a_hi + off) = h;
      *reinterpret_cast<__nv_bfloat16*>(a_lo + off) = l;
    }
  } else {
    // logical A[m=c][k=pos] = A[c*K + pos]
    for (int idx = tid; idx < 128 * p.K; idx += 128) {
      const int c = idx / p.K, pos = idx % p.K;
      const float v = A[c * p.K + pos];
      const __nv_bfloat16 h = __float2bfloat16_rn(v);
      const __nv_bfloat16 l = __float2bfloat16_rn(v - __bfloat162float(h));
      const int off = (c / 8) * PSLOTS * 16 + (pos + p.shift) * 16 + (c % 8) * 2;
      *reinterpret_cast<__nv_bfloat16*>(x_hi + off) = h;
      *reinterpret_cast<__nv_bfloat16*>(x_lo + off) = l;
    }
  }
  for (int idx = tid; idx < Npad * p.K; idx += 128) {
    const int n = idx / p.K, k = idx % p.K;
    const float v = B[n * p.K + k];
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    const __nv_bfloat16 l = __float2bfloat16_rn(v - __bfloat162float(h));
    const int off = (k / 8) * Npad * 16 + n * 16 + (k % 8) * 2;
    *reinterpret_cast<__nv_bfloat16*>(b_hi + off) = h;
    *reinterpret_cast<__nv_bfloat16*>(b_lo + off) = l;
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base;
  if (tid == 0) {
    const uint32_t amaj = p.mode >= 2 ? 1u : 0u;
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (amaj << 15) | ((uint32_t)(p.N >> 3) << 17) | ((128u >> 4) << 24);
    for (int ks = 0; ks < p.K / 16; ++ks) {
      uint64_t dah, dal;
      if (p.mode < 2) {
        const uint32_t lbo = SLOTS * 16, sbo = 128;
        const uint32_t off = ks * 2 * lbo + p.shift * 16;
        dah = p.mode == 0 ? make_desc(smem_u32(a_hi) + off, lbo, sbo) : make_desc(smem_u32(a_hi) + off, sbo, lbo);
        dal = p.mode == 0 ? make_desc(smem_u32(a_lo) + off, lbo, sbo) : make_desc(smem_u32(a_lo) + off, sbo, lbo);
      } else {
        const uint32_t grp_mn = PSLOTS * 16, grp_k = 128;   // stride between 8-c groups / between 8-pos groups
        const uint32_t off = ks * 2 * grp_k + p.shift * 16;
        dah = p.mode == 2 ? make_desc(smem_u32(x_hi) + off, grp_k, grp_mn) : make_desc(smem_u32(x_hi) + off, grp_mn, grp_k);
        dal = p.mode == 2 ? make_desc(smem_u32(x_lo) + off, grp_k, grp_mn) : make_desc(smem_u32(x_lo) + off, grp_mn, grp_k);
      }
      const uint32_t lbo_b = Npad * 16;
      const uint64_t dbh = make_desc(smem_u32(b_hi) + ks * 2 * lbo_b, lbo_b, 128);
      const uint64_t dbl = make_desc(smem_u32(b_lo) + ks * 2 * lbo_b, lbo_b, 128);
      mma_bf16(tmem, dah, dbh, idesc, ks > 0);
      mma_bf16(tmem, dal, dbh, idesc, 1);
      mma_bf16(tmem, dah, dbl, idesc, 1);
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&bar)) : "memory");
  }
  // wait for the MMAs
  {
    uint32_t done = 0;
    while (!done) {
      asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                   : "=r"(done) : "r"(smem_u32(&bar)), "r"(0) : "memory");
    }
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // epilogue: warp w reads TMEM lanes 32w..32w+31, 8 columns at a time
  for (int c0 = 0; c0 < p.N; c0 += 8) {
    uint32_t r[8];
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + c0;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 8; ++j) out[tid * p.N + c0 + j] = __uint_as_float(r[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(256));
}

int main() {
  const int K = 64;
  std::vector<float> A(128 * K), B(256 * K);
  srand(1);
  for (auto& v : A) v = (float)rand() / RAND_MAX - 0.5f;
  for (auto& v : B) v = (float)rand() / RAND_MAX - 0.5f;
  float *dA, *dB, *dO;
  cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dO, 128 * 256 * 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
  const int smem = 200 * 1024;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int Ns[] = {16, 32, 48, 96, 192};
  for (int mode = 0; mode < 4; ++mode)
    for (int shift : {0, 1, 3, 66})
      for (int N : Ns) {
        if (N != 16 && (shift == 3)) continue;
        Params p{mode, shift, N, K};
        cudaMemset(dO, 0, 128 * 256 * 4);
        probe_kernel<<<1, 128, smem>>>(dA, dB, dO, p);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("mode %d shift %d N %d: CUDA error %s\n", mode, shift, N, cudaGetErrorString(e)); return 1; }
        std::vector<float> O(128 * N);
        cudaMemcpy(O.data(), dO, O.size() * 4, cudaMemcpyDeviceToHost);
        double maxerr = 0, maxref = 0;
        for (int m = 0; m < 128; ++m)
          for (int n = 0; n < N; ++n) {
            double ref = 0;
            for (int k = 0; k < K; ++k) ref += (double)A[m * K + k] * B[n * K + k];
            maxerr = fmax(maxerr, fabs(ref - O[m * N + n]));
            maxref = fmax(maxref, fabs(ref));
          }
        printf("mode %d shift %2d N %3d: max|err| %.3e (max|ref| %.3f) %s\n", mode, shift, N, maxerr, maxref,
               maxerr < 1e-4 * maxref ? "OK" : "MISMATCH");
      }
  return 0;
}
