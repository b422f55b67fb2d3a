// Tile geometry shared by the convolution kernels.
#pragma once
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
#endif
}  // namespace mg
