// Thin inline-PTX layer over the sm_100a tensor-core path: tcgen05.mma (UMMA) with shared-memory
// operand descriptors, TMEM allocation / loads, mbarriers and the proxy fences they need.
// No CUTLASS: only the PTX forms, with the descriptor bit layouts written out.
//
// Shared-memory operand layout used throughout this library: the *canonical no-swizzle* layout, in
// units of 16 bytes (= 8 bf16):
//     chunk(c, r)  at  base + (c * rows_pitch + r) * 16 B          c = 8-element chunk index along the
//                                                                  contiguous ("leading") tensor dim,
//                                                                  r = index along the other dim
//   * K-major operand  (MMA rows = r, reduction K = 8c..8c+7):  8-row core matrices of 128 contiguous
//       bytes; SBO (next 8 rows) = 128 B; LBO (next 8 K) = rows_pitch * 16 B.
//   * MN-major operand (MMA rows = 8c..8c+7, reduction K = r):  8-K core matrices of 128 contiguous bytes;
//       LBO (next 8 K) = 128 B; SBO (next 8 rows) = rows_pitch * 16 B.
// i.e. ONE staging layout [channel chunk][pixel][8 channels] serves the forward / data-gradient
// convolutions (K = channels) and the weight-gradient convolution (K = pixels).  Because rows sit at a
// uniform 16-byte pitch, an operand may start at ANY row: the 3x3 taps of a convolution are just nine
// different start addresses into one staged halo tile.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

namespace mg {
namespace umma {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier ---------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("{\n.reg .b64 st;\nmbarrier.arrive.shared::cta.b64 st, [%0];\n}" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n.reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {}
}

// 16-byte asynchronous global -> shared copy (LDGSTS); src_bytes = 0 zero-fills the destination (padding)
__device__ __forceinline__ void cp_async16(uint32_t dst_smem, const void* src, uint32_t src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst_smem), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async16_full(uint32_t dst_smem, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst_smem), "l"(src) : "memory");
}
// arrive on `bar` (counted in its expected arrivals) once all prior cp.async of this thread have landed
__device__ __forceinline__ void cp_async_arrive(uint64_t* bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// generic-proxy smem writes (st.shared) -> visible to the async proxy (tensor core operand reads)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- TMEM -------------------------------------------------------------------------------------
// one full warp; ncols power of two >= 32; writes the TMEM base address to *dst_smem
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// 32 lanes x 16 consecutive fp32 columns: thread i of the warp receives TMEM lane (lane_base + i)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// ---- descriptors --------------------------------------------------------------------------------
// shared-memory matrix descriptor, SWIZZLE_NONE, sm_100 version bits = 1
//   [0,14) start >> 4 | [16,30) leading byte offset >> 4 | [32,46) stride byte offset >> 4 | [46,48) = 1
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
// instruction descriptor for kind::f16, bf16 x bf16 -> fp32, M = 128 (cta_group::1)
//   [4,6) D fmt (1 = f32) | [7,10) A fmt (1 = bf16) | [10,13) B fmt | 15 A major (1 = MN) | 16 B major |
//   [17,23) N >> 3 | [24,29) M >> 4
__host__ __device__ __forceinline__ uint32_t instr_desc_bf16(int n, bool a_mn_major, bool b_mn_major, int m = 128) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((a_mn_major ? 1u : 0u) << 15) | ((b_mn_major ? 1u : 0u) << 16) |
           ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem]; one thread issues
__device__ __forceinline__ void mma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n.reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
// same with the A operand read from TMEM (M rows = lanes, K along columns, two bf16 per 32-bit column)
__device__ __forceinline__ void mma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n.reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
// registers -> TMEM: thread i of the warp writes 8 consecutive 32-bit columns of lane (lane_base + i)
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
// thread i, registers (r0, r1, r2, r3) -> TMEM (lane i/4, col i%4), (lane i/4 + 8, col i%4), (lane i/4, col 4 + i%4),
// (lane i/4 + 8, col 4 + i%4) relative to taddr: the fragment layout of ldmatrix.x4(.trans) / mma.sync operands
__device__ __forceinline__ void tmem_st_16x128b_x2(uint32_t taddr, uint32_t r0, uint32_t r1, uint32_t r2, uint32_t r3) {
    asm volatile("tcgen05.st.sync.aligned.16x128b.x2.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(r0), "r"(r1), "r"(r2), "r"(r3) : "memory");
}
// four 8 x 8 b16 matrices (lane l supplies the 16-byte row l%8 of matrix l/8), transposed on the way: thread i receives,
// per matrix, elements [2(i%4)][i/4] (low half) and [2(i%4)+1][i/4] (high half)
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t saddr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(saddr) : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// arrive on `bar` once every previously issued MMA of this thread has completed (implies fence::before_thread_sync)
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}


// ---- warp-convergent issue -------------------------------------------------------------------------------------------
// tcgen05.mma / tcgen05.commit take their operands from UNIFORM registers.  Issued inside `if (lane == 0)` the operands
// are per-thread registers of a diverged warp and ptxas wraps every instruction in a serialising ELECT / R2UR /
// BRA.U.ANY loop: ~200 cycles per MMA per issuing thread (measured, scripts/probe_mma_pace.py), whatever the shape.  The
// variants below are executed by ALL 32 lanes of the issuing warp with warp-uniform arguments; elect.sync picks the
// issuing lane (the same one every time for a full mask), and the instruction goes out at the tensor pipe's own pace.
__device__ __forceinline__ void mma_bf16_warp(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n.reg .pred p, e;\n"
        "elect.sync _|e, 0xffffffff;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mma_bf16_ts_warp(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n.reg .pred p, e;\n"
        "elect.sync _|e, 0xffffffff;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
// true in exactly one lane of a converged warp (the same lane every time); everything inside `if (elect_one()) { ... }`
// is single-thread code whose R2UR moves need no serialising loop
__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n.reg .pred e;\n"
        "elect.sync _|e, 0xffffffff;\n"
        "@e mov.u32 %0, 1;\n}"
        : "+r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void mma_commit_warp(uint64_t* bar) {
    asm volatile(
        "{\n.reg .pred e;\n"
        "elect.sync _|e, 0xffffffff;\n"
        "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n}"
        ::"r"(smem_u32(bar)) : "memory");
}

}  // namespace umma
}  // namespace mg
