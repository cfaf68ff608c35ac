// tcgen05 / TMEM / mbarrier / bulk-copy wrappers shared by the implicit-GEMM conv kernels (sm_100a).
// Conventions validated on hardware by tests/probe/umma_probe.cu:
//  * K-major operand, no swizzle: element (row r, k) at  (k/8)*LBO + r*16 + (k%8)*2  bytes
//    (rows at a uniform 16-byte pitch => SBO = 128); a descriptor whose start address is moved by
//    s*16 bytes addresses rows r+s ("shifted tap": one staged tile serves every filter tap);
//  * MN-major operand, no swizzle: element (m, k) at (m/8)*SBO + k*16 + (m%8)*2 with LBO = 128;
//  * fp32 data is split as hi = bf16(v), lo = bf16(v - hi) and a product is hi*hi + lo*hi + hi*lo
//    accumulated in fp32 TMEM (~3e-6 relative; single-pass bf16 or tf32 would miss the 1e-3
//    gradient tolerance of the parity contract).
#pragma once
#include <cuda_bf16.h>
#include "common.cuh"

namespace lshm {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}

// descriptor = {lo: start address (>>4) + LBO, hi: SBO + version}; moving the start address by `off16`
// 16-byte units is one 32-bit add on the low word (shared-memory addresses stay below the 14-bit field)
__device__ __forceinline__ uint64_t desc_off(uint64_t d, uint32_t off16) {
  uint32_t lo, hi;
  asm("mov.b64 {%0,%1}, %2;" : "=r"(lo), "=r"(hi) : "l"(d));
  lo += off16;
  uint64_t r;
  asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "r"(lo), "r"(hi));
  return r;
}

// instruction descriptor: bf16 x bf16 -> fp32, M = 128
__host__ __device__ __forceinline__ uint32_t make_idesc(int n, int a_mn_major, int b_mn_major, int m = 128) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
         ((uint32_t)(n >> 3) << 17) | (((uint32_t)m >> 4) << 24);
}

__device__ __forceinline__ void mma_bf16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n"
               :: "r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}

// Same instruction issued from warp-uniform code: every lane of the (converged) warp executes the call,
// one elected lane issues.  The operands then stay in uniform registers; under `if (lane == 0)` the
// compiler wraps every tcgen05.mma in a vote/broadcast "waterfall" (~16 instructions per MMA).
__device__ __forceinline__ uint32_t elect_one() {          // 1 in exactly one lane of the converged warp
  uint32_t r;
  asm volatile("{\n.reg .pred q;\nelect.sync _|q, 0xffffffff;\nselp.u32 %0, 1, 0, q;\n}\n" : "=r"(r));
  return r;
}
__device__ __forceinline__ void mma_bf16_warp(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc,
                                              uint32_t leader) {
  asm volatile("{\n.reg .pred p, q;\nsetp.ne.b32 p, %4, 0;\nsetp.ne.b32 q, %5, 0;\n"
               "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n"
               :: "r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc), "r"(leader) : "memory");
}
__device__ __forceinline__ void mma_split3_warp(uint32_t tmem_d, uint64_t a_hi, uint64_t a_lo, uint64_t b_hi, uint64_t b_lo,
                                                uint32_t idesc, uint32_t acc, uint32_t leader) {
  mma_bf16_warp(tmem_d, a_hi, b_hi, idesc, acc, leader);
  mma_bf16_warp(tmem_d, a_lo, b_hi, idesc, 1u, leader);
  mma_bf16_warp(tmem_d, a_hi, b_lo, idesc, 1u, leader);
}
__device__ __forceinline__ void commit_warp(uint64_t* bar, uint32_t leader) {   // warp-uniform call, the leader commits
  asm volatile("{\n.reg .pred q;\nsetp.ne.b32 q, %1, 0;\n"
               "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n}\n"
               :: "r"(smem_u32(bar)), "r"(leader) : "memory");
}

// D += A*B with the 3-term bf16 split; *_hi / *_lo descriptors address the two halves
__device__ __forceinline__ void mma_split3(uint32_t tmem_d, uint64_t a_hi, uint64_t a_lo, uint64_t b_hi, uint64_t b_lo,
                                           uint32_t idesc, uint32_t acc) {
  mma_bf16(tmem_d, a_hi, b_hi, idesc, acc);
  mma_bf16(tmem_d, a_lo, b_hi, idesc, 1u);
  mma_bf16(tmem_d, a_hi, b_lo, idesc, 1u);
}

__device__ __forceinline__ void commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t* dst, uint32_t ncols) {   // one full warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(dst)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {  // the same warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                 "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, float (&v)[4]) {
  uint32_t r[4];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 4; ++i) v[i] = __uint_as_float(r[i]);
}

// ---- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// Waiting warps must not steal issue slots from the producer warps (an ncu capture of the first
// version showed 55% of all executed instructions were try_wait/branch spins): the retry path backs
// off with a short nanosleep.  (A suspend-time hint on try_wait made every hand-off slower.)
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, uint32_t backoff_ns = 32) {
  uint32_t done = 0;
  const uint32_t addr = smem_u32(bar);
  while (true) {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    if (done) break;
    __nanosleep(backoff_ns);
    // exponential back-off: with a fixed 20 ns the retry loops were still 30-45% of all executed
    // instructions of the conv kernels, which are issue-bound (profiles/r1_ncu_conv_all.md)
    backoff_ns = backoff_ns < 256 ? backoff_ns * 2 : backoff_ns;
  }
}
// position in a ring of `n` pipeline stages: stage index and phase bit, advanced without the
// division/modulo by a run-time stage count (~40 instructions per hand-off in the first version)
struct Ring {
  int s; uint32_t ph;
  __device__ __forceinline__ void next(int n) { if (++s == n) { s = 0; ph ^= 1u; } }
};

// 1-D bulk copy global -> shared, completion counted on `bar` (bytes multiple of 16, 16 B aligned)
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// 16-byte asynchronous copy global -> shared (LDGSTS, L1 bypass); src_bytes = 0 zero-fills
__device__ __forceinline__ void cp_async16(void* dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" :: "r"(smem_u32(dst)), "l"(src), "r"(src_bytes) : "memory");
}
// arrive on `bar` once all cp.async issued so far by this thread have landed (count pre-set at init)
__device__ __forceinline__ void cp_async_arrive(uint64_t* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
}

// ---- fp32 -> (hi, lo) bf16 split, 8 values -> two 16-byte chunks
__device__ __forceinline__ void split8(const float (&v)[8], uint4& hi, uint4& lo) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    // packed converts: 2 cvt + 2 unpack + 2 sub per pair of values
    const __nv_bfloat162 h2 = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    const uint32_t hp = *reinterpret_cast<const uint32_t*>(&h2);
    const float h0 = __uint_as_float(hp << 16), h1 = __uint_as_float(hp & 0xffff0000u);
    const __nv_bfloat162 l2 = __floats2bfloat162_rn(v[2 * i] - h0, v[2 * i + 1] - h1);
    h[i] = hp;
    l[i] = *reinterpret_cast<const uint32_t*>(&l2);
  }
  hi = make_uint4(h[0], h[1], h[2], h[3]);
  lo = make_uint4(l[0], l[1], l[2], l[3]);
}

}  // namespace tc
}  // namespace lshm
