// Tile geometry shared by the convolution kernels.
#pragma once
namespace mg {
constexpr int kTileH = 16;                       // output rows per tile   (= 16 eight-row core-matrix groups)
constexpr int kTileW = 8;                        // output cols per tile   (= one core-matrix group)
constexpr int kHaloW = kTileW + 2;               // 10
constexpr int kHaloH = kTileH + 2;               // 18
constexpr int kHaloPos = kHaloW * kHaloH;        // 180 halo positions
constexpr int kHaloPitch = 186;                  // positions per channel chunk in smem (== 2 mod 8: spreads banks)
}  // namespace mg
