/* musicgan_b200 -- TEST-ONLY probes of the tcgen05 / TMEM conventions the convolution kernels rely on.
 * Built into libmusicgan_b200_debug.so (not into the product library); used by tests/test_umma_probe_gpu.py and
 * scripts/probe_tmem_store.py. */
#ifndef MUSICGAN_B200_DEBUG_H
#define MUSICGAN_B200_DEBUG_H
#include "musicgan_b200.h"
#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------------------------------
 * Test-only probe of the tcgen05 / TMEM conventions the convolution kernels rely on (one tile).
 * mode 0: A [Ra][K], B [N][K] bf16 (K contiguous);  D[m][n] = sum_k A[row_off + (m/8)*grp_rows + m%8][k] * B[n][k]
 * mode 1: A [K][128], B [K][N] bf16 (M / N contiguous);  D[m][n] = sum_k A[k][m] * B[k][n]
 * D [128][N] fp32.
 * ---------------------------------------------------------------------------------------- */
int mg_debug_umma_gemm(const void* A, const void* B, float* D, int K, int N, int mode, int Ra, int row_off,
                       int grp_rows, mgStream stream);
/* Test-only probe of the tcgen05.st shapes used by the weight-gradient kernel's operand staging: one store of
 * shape 0: 16x64b.x1, 1: 16x128b.x1, 2: 16x128b.x2, 3: 16x256b.x1 by warp 0 with register values
 * 0x1000 | thread << 4 | (register index + 1); out [128 lanes][32 columns] uint32 = the TMEM block afterwards. */
int mg_debug_tmem_store(uint32_t* out, int shape, int lane_off, int col_off, mgStream stream);

/* Pacing probe: clock64 ticks one thread needs to issue AND complete a chain of n_mma M128 x N x K16 bf16 MMAs rotating
 * over n_acc accumulators (1 = each depends on the previous through D).  mode 0: operands in shared memory, 8-row groups
 * 128 B apart; 1: 160 B apart (the convolutions' halo rows); 2: A operand in TMEM. */
int mg_debug_mma_pace(long long* cycles_dev, int N, int n_mma, int n_acc, int mode, mgStream stream);
/* The single-thread-region chain on `ctas` CTAs at once (128 TMEM columns, 49 KB of shared memory each: up to four share
 * an SM); cycles_dev[cta] = that CTA's ticks. */
int mg_debug_mma_pace_grid(long long* cycles_dev, int N, int n_mma, int ctas, mgStream stream);

/* Contention probe: a chain of n_mma MMAs (N columns) while four warps drain `ld_cols` other accumulator columns in a loop.
 * out_dev[0] = ticks of the chain, out_dev[1] = 16-column loads completed by each draining warp meanwhile. */
int mg_debug_mma_vs_drain(long long* out_dev, int N, int n_mma, int ld_cols, mgStream stream);

#ifdef __cplusplus
}
#endif
#endif /* MUSICGAN_B200_DEBUG_H */
