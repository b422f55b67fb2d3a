// Tile geometry shared by the convolution kernels.
#pragma once
#include <cuda_bf16.h>
namespace mg {
constexpr int kTileH = 16;                       // output rows per tile   (= 16 eight-row core-matrix groups)
constexpr int kTileW = 8;                        // output cols per tile   (= one core-matrix group)
constexpr int kHaloW = kTileW + 2;               // 10
constexpr int kHaloH = kTileH + 2;               // 18
constexpr int kHaloPos = kHaloW * kHaloH;        // 180 halo positions
constexpr int kHaloPitch = 186;                  // positions per channel chunk in smem (== 2 mod 8: spreads banks)
constexpr int kMaxStages = 12;                   // halo slots in flight per CTA (memory-level parallelism: ~60 KB/SM)

// i -> (pos, c) with i = pos * nch + c, via a host-computed reciprocal (exact for i * nch < 2^20)
struct ItemDiv { unsigned magic; int nch; };
inline ItemDiv make_item_div(int nch) { ItemDiv d; d.magic = ((1u << 20) + nch - 1) / nch; d.nch = nch; return d; }

// n / d for 0 <= n < 2^20 and 1 <= d < 2^20 with a host-computed 41-bit reciprocal: q = (n * M) >> 40
struct FastDiv { unsigned long long m; };
inline FastDiv make_fast_div(int d) { FastDiv f; f.m = ((1ull << 40) / (unsigned long long)d) + 1ull; return f; }
#if defined(__CUDACC__)
__device__ __forceinline__ int fast_div(int n, FastDiv f) { return (int)(((unsigned long long)(unsigned)n * f.m) >> 40); }

// Weight packing of the split-operand kernel (conv_split.cu):
// fp32 [Cout][Cin][3][3] -> split bf16 [K/16][tap][2 chunks][part][N][8]   (N, K = GEMM channel counts); `wparts` = 2:
// w = hi + lo (16 significand bits), 3: w = hi + mid + lo (24 bits, the fp32 weight exactly).  The parts of a chunk sit
// side by side so that ONE MMA against [hi | mid | lo] (GEMM N = parts * slice) computes all products of an activation half.
//   transpose_flip = 0: B[n = co][k = ci] of tap (ky,kx) = w[co][ci][ky][kx]
//   transpose_flip = 1: data gradient, B[n = ci][k = co] = w[co][ci][2-ky][2-kx]
__device__ __forceinline__ void pack_weights_split_range(const float* __restrict__ w, int Cout, int Cin, int transpose_flip, int wparts,
                                                         __nv_bfloat16* __restrict__ out, int first, int stride) {
    const int N = transpose_flip ? Cin : Cout, K = transpose_flip ? Cout : Cin;
    const int total = 9 * N * K;                       // (hi, lo) pairs
    for (int i = first; i < total; i += stride) {
        int r = i;
        const int e = r & 7; r >>= 3;
        const int n = r % N; r /= N;
        const int chunk = r & 1; r >>= 1;
        const int tap = r % 9, cg = r / 9;
        const int k = cg * 16 + chunk * 8 + e;
        const int ky = tap / 3, kx = tap % 3;
        const float v = transpose_flip ? w[(((size_t)k * Cin + n) * 3 + (2 - ky)) * 3 + (2 - kx)]
                                       : w[(((size_t)n * Cin + k) * 3 + ky) * 3 + kx];
        const __nv_bfloat16 hi = __float2bfloat16_rn(v);
        const float r1 = v - __bfloat162float(hi);
        const __nv_bfloat16 mid = __float2bfloat16_rn(r1), lo = __float2bfloat16_rn(r1 - __bfloat162float(mid));
        // destination: [cg][tap][chunk][part][n][e]
        const size_t base = ((size_t)(cg * 9 + tap) * 2 + chunk) * wparts * N * 8;
        out[base + ((size_t)0 * N + n) * 8 + e] = hi;
        out[base + ((size_t)1 * N + n) * 8 + e] = mid;
        if (wparts == 3) out[base + ((size_t)2 * N + n) * 8 + e] = lo;
    }
}
#endif
}  // namespace mg
