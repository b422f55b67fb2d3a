// Tile geometry shared by the convolution kernels.
#pragma once
namespace mg {
constexpr int kTileH = 16;                       // output rows per tile   (= 16 eight-row core-matrix groups)
constexpr int kTileW = 8;                        // output cols per tile   (= one core-matrix group)
constexpr int kHaloW = kTileW + 2;               // 10
constexpr int kHaloH = kTileH + 2;               // 18
constexpr int kHaloPos = kHaloW * kHaloH;        // 180 halo positions
constexpr int kHaloPitch = 186;                  // positions per channel chunk in smem (== 2 mod 8: spreads banks)
constexpr int kMaxStages = 4;                    // halo slots in flight per CTA

// i -> (pos, c) with i = pos * nch + c, via a host-computed reciprocal (exact for i * nch < 2^20)
struct ItemDiv { unsigned magic; int nch; };
inline ItemDiv make_item_div(int nch) { ItemDiv d; d.magic = ((1u << 20) + nch - 1) / nch; d.nch = nch; return d; }
}  // namespace mg
