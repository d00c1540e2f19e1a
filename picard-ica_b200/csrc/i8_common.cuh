// =====================================================================================================
// i8_common.cuh -- what the two INT8 tensor-core passes (i8_loss.cu, i8_grad.cu) share: the error-free splitting of an
// f64 operand into balanced radix-256 digits, the tcgen05 / TMEM / bulk-copy PTX wrappers and the UMMA descriptors.
//
// Splitting (Ozaki-type, exact integer accumulation).  A value v with a power-of-two bound 1.008 |v| < 2^e is rounded to the
// fixed-point integer I = rint(v 2^(8S-1-e)), |I| < 0.993 2^(8S-1), and written in BALANCED radix-256 digits
//     I = sum_{p=0}^{S-1} q_p 256^(S-1-p),   q_p in [-128, 127]   (p = 0 most significant),
// i.e. v ~= 2^e sum_p q_p 2^(-8p-7).  The digits are the bytes of U = I + sum_p 128 256^p with their top bit flipped
// (offset binary -> two's complement): one 64-bit add and byte extraction, no carry loop.  For two operands
//     a b ~= 2^(ea+eb-14) sum_{p,q} a_p b_q 2^(-8(p+q)),
// and the products with p + q = d <= S - 1 are INT8 GEMMs whose sums are exact in the s32 accumulator of their LEVEL d.  With
// S = 6 that is 21 products (the sign-magnitude 7-bit digits of round 1 needed S = 7 and 28 products for the same width:
// a signed byte carries 8 bits only when the digit set is balanced).  Dropped: the products with p + q >= S, below
// (S-1) 2^(14-8(S+2)+16) = 5 2^-50 of the two bounds' product each, and -- because balanced digits have signs independent of
// the value's sign -- zero-mean, so their sum over a contraction grows like sqrt(K), not K (tools/ozaki_numerics.py).
// Level combination for S = 6 (exact in 64-bit integers):
//     sum_d 2^(-8d) L_d = 2^-16 (hi + 2^-24 lo),  hi = L0 2^16 + L1 2^8 + L2,  lo = L3 2^16 + L4 2^8 + L5.
// =====================================================================================================
#pragma once
#include <cmath>
#include <cstdint>

#include "i8_split.h"
#include "pass.cuh"

namespace picard {
namespace i8 {

#ifdef __CUDACC__
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// shared-memory matrix descriptor: K-major, SWIZZLE_128B, rows of 128 bytes; 8-row groups 1024 bytes apart (SBO); LBO unused
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version of sm_100
  d |= (uint64_t)2 << 61;  // SWIZZLE_128B
  return d;
}
// instruction descriptor: D = s32, A = B = signed int8, both K-major, M = 128, N = n
__host__ __device__ constexpr uint32_t make_idesc(int n) {
  return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
// D[tmem_d] (+)= A[tmem_a] * B[smem descriptor]: A (128 rows x 32 K-bytes) in tensor memory, row = lane, 8 columns of 4 bytes
__device__ __forceinline__ void umma_i8_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(db), "r"(idesc), "r"(accumulate), "r"(0u)
      : "memory");
}
// D[tmem_d] (+)= A[smem descriptor] * B[smem descriptor]
__device__ __forceinline__ void umma_i8_ss(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(accumulate), "r"(0u)
      : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]),
               "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st4(uint32_t taddr, const uint32_t (&v)[4]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src),
               "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// 4 / 8 consecutive TMEM columns of this thread's lane
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, int32_t* v) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]) : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, int32_t* v) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, int32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
        "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc512(uint32_t* slot) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(slot)) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc512(uint32_t tmem) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}
// v 2^k on the ALU (v = 0 or a normal double far from the exponent limits): keeps the FP64 pipe for the densities
__device__ __forceinline__ double scale_pow2(double v, int k) {
  const int h = __double2hiint(v), l = __double2loint(v);
  return (((h << 1) | l) != 0) ? __hiloint2double(h + (k << 20), l) : v;
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// TMA: 2-D tiled tensor store shared -> global (SASS: UTMASTG), bulk-group completion
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* tmap, const void* src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(tmap), "r"(c0), "r"(c1),
               "r"(smem_u32(src))
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
#endif  // __CUDACC__

}  // namespace i8
}  // namespace picard
