// Probe / unit-test entry for the UMMA conventions in umma.cuh (descriptor bit layout, canonical
// no-swizzle staging, row-offset starts, MN-major operands, TMEM lane mapping).  One CTA, one tile.
// Not on any product path; called only by tests/test_umma_probe_gpu.py.
#include "common.cuh"
#include "../../include/musicgan_b200_debug.h"
#include "umma.cuh"

namespace mg {
using namespace umma;

// stage G[rows][cols] (cols contiguous, bf16) as S[(c * pitch + r)] = G[r][8c .. 8c+7]  (16-byte chunks)
__device__ void stage_chunks(const __nv_bfloat16* __restrict__ g, int rows, int cols, uint4* s, int pitch) {
    const int chunks = cols / 8;
    for (int i = threadIdx.x; i < rows * chunks; i += blockDim.x) {
        const int r = i / chunks, c = i % chunks;
        s[c * pitch + r] = *reinterpret_cast<const uint4*>(g + (size_t)r * cols + 8 * c);
    }
}

// mode 0: A [Ra][K] K-major rows, B [N][K] K-major rows.  D[m][n] = sum_k A[row_off + (m/8)*grp_rows + m%8][k] B[n][k]
// mode 1: A [K][128] MN-major, B [K][N] MN-major.          D[m][n] = sum_k A[k][m] B[k][n]
__global__ void __launch_bounds__(128)
k_debug_umma(const __nv_bfloat16* __restrict__ A, const __nv_bfloat16* __restrict__ B, float* __restrict__ D,
             int K, int N, int mode, int Ra, int row_off, int grp_rows) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint4* sA; uint4* sB; int pA, pB;
    // modes 2 / 3: A [128][K] row major goes to TMEM (thread m = row m, pairs of K packed per 32-bit column; mode 3 swaps
    // the halves of each pair), B [K][N] MN-major in smem as in mode 1
    if (mode == 0) { pA = Ra + 1; pB = N + 1; } else { pA = K + 1; pB = K + 1; }
    sA = reinterpret_cast<uint4*>(smem);
    sB = sA + (mode == 0 ? (K / 8) * pA : (128 / 8) * pA);
    if (mode == 0) { stage_chunks(A, Ra, K, sA, pA); stage_chunks(B, N, K, sB, pB); }
    else if (mode == 1) { stage_chunks(A, K, 128, sA, pA); stage_chunks(B, K, N, sB, pB); }
    else stage_chunks(B, K, N, sB, pB);
    if (warp == 0) tmem_alloc(&tmem_base, 512);
    if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = tmem_base;
    const uint32_t tm_a = tm + 256;           // A operand columns (modes 2 / 3) after the accumulator
    if (mode >= 2) {
        const int m = tid;
        for (int k0 = 0; k0 < K; k0 += 16) {
            uint32_t r[8];
            for (int j = 0; j < 8; ++j) {
                const uint32_t lo = *reinterpret_cast<const uint16_t*>(A + (size_t)m * K + k0 + 2 * j);
                const uint32_t hi = *reinterpret_cast<const uint16_t*>(A + (size_t)m * K + k0 + 2 * j + 1);
                r[j] = mode == 2 ? (lo | (hi << 16)) : (hi | (lo << 16));
            }
            tmem_st8(tm_a + ((uint32_t)(warp * 32) << 16) + (k0 >> 1), r);
        }
        tmem_wait_st();
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
    }
    if (tid == 0) {
        const uint32_t idesc = instr_desc_bf16(N, mode == 1, mode == 1);
        (void)idesc;
        const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
        for (int kk = 0; kk < K / 16; ++kk) {
            uint64_t da, db;
            if (mode >= 2) {
                db = smem_desc(b0 + (kk * 16) * 16, 128, pB * 16);
                mma_bf16_ts(tm, tm_a + kk * 8, db, instr_desc_bf16(N, false, true), kk > 0 ? 1u : 0u);
                continue;
            }
            if (mode == 0) {
                da = smem_desc(a0 + (row_off + kk * 2 * pA) * 16, pA * 16, grp_rows * 16);
                db = smem_desc(b0 + (kk * 2 * pB) * 16, pB * 16, 128);
            } else {
                da = smem_desc(a0 + (kk * 16) * 16, 128, pA * 16);
                db = smem_desc(b0 + (kk * 16) * 16, 128, pB * 16);
            }
            mma_bf16(tm, da, db, idesc, kk > 0 ? 1u : 0u);
        }
        mma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    tc_fence_after();
    for (int n0 = 0; n0 < N; n0 += 16) {
        float v[16];
        tmem_ld16(tm + ((uint32_t)(warp * 32) << 16) + n0, v);
        tmem_wait_ld();
        const int m = warp * 32 + lane;
        for (int j = 0; j < 16; ++j) D[(size_t)m * N + n0 + j] = v[j];
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tm, 512);
}
}  // namespace mg

extern "C" int mg_debug_umma_gemm(const void* A, const void* B, float* D, int K, int N, int mode, int Ra, int row_off,
                                  int grp_rows, mgStream stream) {
    using namespace mg;
    if (!A || !B || !D || K % 16 || N % 16 || N < 16 || N > 256) return MG_ERR_BAD_ARG;
    size_t bytes;
    if (mode == 0) bytes = ((size_t)(K / 8) * (Ra + 1) + (size_t)(K / 8) * (N + 1)) * 16;
    else bytes = ((size_t)16 * (K + 1) + (size_t)(N / 8) * (K + 1)) * 16;
    if (bytes > 200 * 1024) return MG_ERR_UNSUPPORTED;
    cudaFuncSetAttribute(k_debug_umma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    k_debug_umma<<<1, 128, bytes, (cudaStream_t)stream>>>((const __nv_bfloat16*)A, (const __nv_bfloat16*)B, D, K, N, mode, Ra,
                                                          row_off, grp_rows);
    return check_launch("k_debug_umma");
}

// ------------------------------------------------------------------------------------------------
// Probe of the tcgen05.st data-path shapes (which thread register lands in which TMEM lane / column):
// 128 threads; the TMEM block [128 lanes][32 columns] is zero-filled, warp 0 executes ONE store of the shape under
// test with register values tag(thread, register index), then every warp dumps its lane quarter with the known
// 32x32b load.  out[lane][col] (uint32), 0 = untouched.
//   shape 0: 16x64b.x1 (1 reg)   1: 16x128b.x1 (2 regs)   2: 16x128b.x2 (4 regs)   3: 16x256b.x1 (4 regs)
//   lane_off / col_off: lane and column fields of the store address
// ------------------------------------------------------------------------------------------------
namespace mg {
__global__ void __launch_bounds__(128)
k_debug_tmem_store(uint32_t* __restrict__ out, int shape, int lane_off, int col_off) {
    using namespace umma;
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) tmem_alloc(&slot, 32);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t base = slot;
    const uint32_t mine = base + ((uint32_t)(warp * 32) << 16);
    uint32_t z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int c = 0; c < 32; c += 8) tmem_st8(mine + c, z);
    tmem_wait_st();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 0) {
        const uint32_t a = base + ((uint32_t)lane_off << 16) + (uint32_t)col_off;
        const uint32_t t = 0x1000u | ((uint32_t)lane << 4);
        if (shape == 0) asm volatile("tcgen05.st.sync.aligned.16x64b.x1.b32 [%0], {%1};" ::"r"(a), "r"(t | 1u) : "memory");
        else if (shape == 1) asm volatile("tcgen05.st.sync.aligned.16x128b.x1.b32 [%0], {%1, %2};" ::"r"(a), "r"(t | 1u), "r"(t | 2u) : "memory");
        else if (shape == 2) asm volatile("tcgen05.st.sync.aligned.16x128b.x2.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(t | 1u), "r"(t | 2u), "r"(t | 3u), "r"(t | 4u) : "memory");
        else asm volatile("tcgen05.st.sync.aligned.16x256b.x1.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(t | 1u), "r"(t | 2u), "r"(t | 3u), "r"(t | 4u) : "memory");
        tmem_wait_st();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    for (int c = 0; c < 32; c += 16) {
        float v[16];
        tmem_ld16(mine + c, v);
        tmem_wait_ld();
        for (int j = 0; j < 16; ++j) out[(warp * 32 + lane) * 32 + c + j] = __float_as_uint(v[j]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(base, 32);
}
}  // namespace mg

extern "C" int mg_debug_tmem_store(uint32_t* out, int shape, int lane_off, int col_off, mgStream stream) {
    using namespace mg;
    if (!out || shape < 0 || shape > 3) return MG_ERR_BAD_ARG;
    k_debug_tmem_store<<<1, 128, 0, (cudaStream_t)stream>>>(out, shape, lane_off, col_off);
    return check_launch("k_debug_tmem_store");
}

// ------------------------------------------------------------------------------------------------
// Pacing probe: how many cycles does ONE elected thread need per M128 x N x K16 bf16 tcgen05.mma when it issues a chain
// of `n_mma` of them (operand contents irrelevant)?  n_acc: the chain rotates over that many accumulators (1 = every
// MMA depends on the previous one through D).  mode 0: A, B from shared memory, canonical no-swizzle, 8-row groups 128 B
// apart (aligned);  1: the same with the groups 160 B apart (the convolution's halo rows);  2: A from TMEM.
// cycles[0] = clock64 ticks from the first issue to the completion of the last MMA.
// ------------------------------------------------------------------------------------------------
namespace mg {
__global__ void __launch_bounds__(128)
k_debug_mma_pace(long long* __restrict__ cycles, int N, int n_mma, int n_acc, int mode, int tmem_cols) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t bar, bar2;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) mbar_init(&bar2, 100000);
    for (int i = tid; i < 48 * 1024 / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    if (warp == 0) tmem_alloc(&tmem_base, tmem_cols);
    if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = tmem_base;
    if (tid == 0 && mode < 4) {
        const uint32_t idesc = instr_desc_bf16(N, false, false);
        const uint32_t a0 = (smem_u32(smem) + 127u) & ~127u, b0 = a0 + 24 * 1024;
        const uint64_t da = smem_desc(a0, 3072u, mode == 1 ? 160u : 128u);
        const uint64_t db = smem_desc(b0, (uint32_t)N * 16u, 128u);
        const uint32_t tm_a = tm + 448;
        const long long t0 = clock64();
        for (int i = 0; i < n_mma; ++i) {
            const uint32_t d = tm + (uint32_t)((i % n_acc) * N);
            if (mode == 2) mma_bf16_ts(d, tm_a, db, idesc, 1u);
            else mma_bf16(d, da, db, idesc, 1u);
        }
        mma_commit(&bar);
        mbar_wait(&bar, 0);
        cycles[0] = clock64() - t0;
    }
    // modes 8 / 9: ONE elect.sync, then the whole chain as single-thread code (9: unrolled by 8, descriptors stepping like
    // the convolution's taps)
    if (warp == 1 && mode >= 8) {
        const uint32_t idesc = instr_desc_bf16(N, false, false);
        const uint32_t a0 = (smem_u32(smem) + 127u) & ~127u, b0 = a0 + 24 * 1024;
        const uint64_t da0 = smem_desc(a0, 3072u, 160u);
        const uint64_t db0 = smem_desc(b0, (uint32_t)N * 16u, 128u);
        const long long t0 = clock64();
        if (elect_one()) {
            if (mode == 11) {      // as mode 8, for many co-resident CTAs (grid > 1): two accumulators
                int acc = 0;
                for (int i = 0; i < n_mma; ++i) {
                    mma_bf16(tm + (uint32_t)(acc * N), da0, db0, idesc, 1u);
                    acc ^= 1;
                }
            } else if (mode == 10) {      // a tcgen05.commit (to a barrier nobody waits on) after every n_acc MMAs, rotating 2 accumulators
                int run = 0, acc = 0;
                for (int i = 0; i < n_mma; ++i) {
                    mma_bf16(tm + (uint32_t)(acc * N), da0, db0, idesc, 1u);
                    if (++run == n_acc) { run = 0; acc ^= 1; mma_commit(&bar2); }
                }
            } else if (mode == 8) {
                int acc = 0;
                for (int i = 0; i < n_mma; ++i) {
                    mma_bf16(tm + (uint32_t)(acc * N), da0, db0, idesc, 1u);
                    if (++acc == n_acc) acc = 0;
                }
            } else {
                for (int i = 0; i < n_mma; i += 8) {
                    uint64_t da = da0, db = db0;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        mma_bf16(tm, da, db, idesc, 1u);
                        da += 1; db += 2;
                    }
                }
            }
            mma_commit(&bar);
        }
        __syncwarp();
        mbar_wait(&bar, 0);
        if ((tid & 31) == 0) cycles[blockIdx.x] = clock64() - t0;
    }
    // modes 4 / 5 / 6: the same chains (as modes 0 / 1 / 2) issued warp-convergently (all lanes of warp 1, elect.sync)
    if (warp == 1 && mode >= 4 && mode < 8) {
        const uint32_t idesc = instr_desc_bf16(N, false, false);
        const uint32_t a0 = (smem_u32(smem) + 127u) & ~127u, b0 = a0 + 24 * 1024;
        const uint64_t da = smem_desc(a0, 3072u, mode == 5 ? 160u : 128u);
        const uint64_t db = smem_desc(b0, (uint32_t)N * 16u, 128u);
        const uint32_t tm_a = tm + 448;
        const long long t0 = clock64();
        int acc = 0;
        for (int i = 0; i < n_mma; ++i) {
            const uint32_t d = tm + (uint32_t)(acc * N);
            if (mode == 6) mma_bf16_ts_warp(d, tm_a, db, idesc, 1u);
            else mma_bf16_warp(d, da, db, idesc, 1u);
            if (++acc == n_acc) acc = 0;
        }
        mma_commit_warp(&bar);
        mbar_wait(&bar, 0);
        if ((tid & 31) == 0) cycles[0] = clock64() - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tm, tmem_cols);
}
}  // namespace mg

extern "C" int mg_debug_mma_pace_grid(long long* cycles, int N, int n_mma, int ctas, mgStream stream) {
    using namespace mg;
    if (!cycles || N % 16 || N < 16 || N > 64 || ctas < 1 || n_mma < 1) return MG_ERR_BAD_ARG;
    cudaFuncSetAttribute(k_debug_mma_pace, cudaFuncAttributeMaxDynamicSharedMemorySize, 49 * 1024);
    k_debug_mma_pace<<<ctas, 128, 49 * 1024, (cudaStream_t)stream>>>(cycles, N, n_mma, 2, 11, 128);
    return check_launch("k_debug_mma_pace");
}

extern "C" int mg_debug_mma_pace(long long* cycles, int N, int n_mma, int n_acc, int mode, mgStream stream) {
    using namespace mg;
    if (!cycles || N % 16 || N < 16 || N > 256 || n_acc < 1 || (mode != 10 && n_acc * N > 448) || n_mma < 1) return MG_ERR_BAD_ARG;
    cudaFuncSetAttribute(k_debug_mma_pace, cudaFuncAttributeMaxDynamicSharedMemorySize, 49 * 1024);
    k_debug_mma_pace<<<1, 128, 49 * 1024, (cudaStream_t)stream>>>(cycles, N, n_mma, n_acc, mode, 512);
    return check_launch("k_debug_mma_pace");
}

// ------------------------------------------------------------------------------------------------
// Do accumulator drains (tcgen05.ld) and MMAs contend?  Warp 1 issues a chain of n_mma M128 x N x K16 MMAs into columns
// [0, N); warps 4..7 (one per TMEM lane quarter) meanwhile read `ld_cols` OTHER columns in a loop (32x32b.x16 loads) until
// the chain has completed.  out[0] = ticks of the chain, out[1] = number of 16-column loads each loader warp completed.
// ------------------------------------------------------------------------------------------------
namespace mg {
__global__ void __launch_bounds__(256)
k_debug_mma_vs_drain(long long* __restrict__ out, int N, int n_mma, int ld_cols) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base;
    __shared__ volatile int done;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < 48 * 1024 / 16; i += 256) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    if (warp == 0) tmem_alloc(&tmem_base, 512);
    if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); done = 0; }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = tmem_base;
    if (warp == 1) {
        const uint32_t idesc = instr_desc_bf16(N, false, false);
        const uint32_t a0 = (smem_u32(smem) + 127u) & ~127u, b0 = a0 + 24 * 1024;
        const uint64_t da0 = smem_desc(a0, 3072u, 160u);
        const uint64_t db0 = smem_desc(b0, (uint32_t)N * 16u, 128u);
        const long long t0 = clock64();
        if (elect_one()) {
            for (int i = 0; i < n_mma; ++i) mma_bf16(tm, da0, db0, idesc, 1u);
            mma_commit(&bar);
        }
        __syncwarp();
        mbar_wait(&bar, 0);
        if (lane == 0) { out[0] = clock64() - t0; done = 1; }
    } else if (warp >= 4 && ld_cols > 0) {
        const uint32_t taddr = tm + 256 + ((uint32_t)((warp & 3) * 32) << 16);
        long long n = 0;
        float sink = 0.0f;
        while (!done) {
            for (int c = 0; c < ld_cols; c += 16) {
                float v[16];
                tmem_ld16(taddr + c, v);
                tmem_wait_ld();
                sink += v[0];
                ++n;
            }
        }
        if (lane == 0 && warp == 4) out[1] = n;
        if (sink == 123.456f) out[2] = 1;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tm, 512);
}
}  // namespace mg

extern "C" int mg_debug_mma_vs_drain(long long* out, int N, int n_mma, int ld_cols, mgStream stream) {
    using namespace mg;
    if (!out || N % 16 || N < 16 || N > 256 || ld_cols < 0 || ld_cols > 256 || (ld_cols & 15) || n_mma < 1) return MG_ERR_BAD_ARG;
    cudaFuncSetAttribute(k_debug_mma_vs_drain, cudaFuncAttributeMaxDynamicSharedMemorySize, 49 * 1024);
    k_debug_mma_vs_drain<<<1, 256, 49 * 1024, (cudaStream_t)stream>>>(out, N, n_mma, ld_cols);
    return check_launch("k_debug_mma_vs_drain");
}
