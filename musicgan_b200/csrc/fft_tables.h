// Host-side construction (double precision, rounded once to fp32) of the constant tables the
// transform kernels use.  Plain C++ so the emulator test can include it too.
#pragma once
#include <cmath>
#include "fft512.cuh"

namespace mg {

inline void build_fft_tables(FftTables* tb) {
    const double two_pi = 6.283185307179586476925286766559;
    for (int r = 0; r < 8; ++r)
        for (int k = 0; k < 8; ++k) {
            const double a = -two_pi * (double)(r * k) / 64.0;
            tb->tw2[r * 8 + k] = make_float2((float)std::cos(a), (float)std::sin(a));
        }
    for (int r = 0; r < 8; ++r)
        for (int j = 0; j < 64; ++j) {
            const double a = -two_pi * (double)(r * j) / 512.0;
            tb->tw3[r * 64 + j] = make_float2((float)std::cos(a), (float)std::sin(a));
        }
}

// W_1024^k = exp(-2 pi i k / 1024), k = 0..511 (real <-> half-complex split twiddles)
inline void build_split_twiddles(float2* w) {
    const double two_pi = 6.283185307179586476925286766559;
    for (int k = 0; k < 512; ++k) {
        const double a = -two_pi * (double)k / 1024.0;
        w[k] = make_float2((float)std::cos(a), (float)std::sin(a));
    }
}

}  // namespace mg
