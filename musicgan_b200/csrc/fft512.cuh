// 512-point complex FFT carried by ONE warp (32 lanes x 16 points), Stockham radix-8 x 3.
//
// The lane-level phases below are plain functions of (lane, registers, exchange buffer) so the
// same source runs (a) inside the sm_100a kernels, where the phases are separated by
// __syncwarp(), and (b) inside tests/emu/fft_emu.cpp, which steps 32 emulated lanes through the
// phases on the host to check the index arithmetic and count shared-memory bank conflicts.
//
// Replaces (for the transform kernels) the library FFT the reference reaches through
// torchaudio.functional.spectrogram / inverse_spectrogram -> torch.stft / torch.istft
// (reference music_gan/audio/functions.py:53-59 and :132-137).
//
// Data layout.  Butterfly j (0..63) of a pass reads logical elements j + 64 r (r = 0..7).
// Lane l owns butterflies j = l (registers v[0..7]) and j = l + 32 (registers v[8..15]).
//   pass 1 (Ns = 1):  no twiddle, writes logical 8 j + r
//   pass 2 (Ns = 8):  twiddle W_64^(r k),  k = j & 7, writes logical 64 (j >> 3) + 8 r + k
//   pass 3 (Ns = 64): twiddle W_512^(r j), result Z[j + 64 r] stays in registers
// so after pass 3 lane l holds Z[l + 32 m], m = 0..15, in slot(m) = (m & 1) * 8 + (m >> 1).
// The two exchanges go through a 512 x float2 warp-private buffer with XOR swizzles that make
// every 64-bit access conflict free per half warp (verified by the emulator).
#pragma once

#if defined(__CUDACC__)
#include <cuda_runtime.h>
#define MG_HD __host__ __device__ __forceinline__
#else
#include <cmath>
#define MG_HD inline
struct float2 { float x, y; };
static inline float2 make_float2(float x, float y) { float2 r; r.x = x; r.y = y; return r; }
#endif

namespace mg {

constexpr int kFftN = 512;          // complex points (= n_fft / 2)
constexpr float kSqrtHalf = 0.70710678118654752440f;

MG_HD float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
MG_HD float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
MG_HD float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
MG_HD float2 cmul_negi(float2 a) { return make_float2(a.y, -a.x); }          // a * (-i)

// In-place forward 8-point DFT, natural order in and out (decimation in frequency).
MG_HD void dft8(float2* v) {
    float2 a0 = cadd(v[0], v[4]), c0 = csub(v[0], v[4]);
    float2 a1 = cadd(v[1], v[5]), c1 = csub(v[1], v[5]);
    float2 a2 = cadd(v[2], v[6]), c2 = csub(v[2], v[6]);
    float2 a3 = cadd(v[3], v[7]), c3 = csub(v[3], v[7]);
    // c_n *= W8^n
    c1 = make_float2((c1.x + c1.y) * kSqrtHalf, (c1.y - c1.x) * kSqrtHalf);
    c2 = cmul_negi(c2);
    c3 = make_float2((c3.y - c3.x) * kSqrtHalf, -(c3.x + c3.y) * kSqrtHalf);
    // 4-point DFT of a -> X0 X2 X4 X6 ; of c -> X1 X3 X5 X7
    float2 e0 = cadd(a0, a2), e1 = csub(a0, a2), o0 = cadd(a1, a3), o1 = cmul_negi(csub(a1, a3));
    float2 f0 = cadd(c0, c2), f1 = csub(c0, c2), p0 = cadd(c1, c3), p1 = cmul_negi(csub(c1, c3));
    v[0] = cadd(e0, o0); v[4] = csub(e0, o0); v[2] = cadd(e1, o1); v[6] = csub(e1, o1);
    v[1] = cadd(f0, p0); v[5] = csub(f0, p0); v[3] = cadd(f1, p1); v[7] = csub(f1, p1);
}

MG_HD int swz1(int i) { return i ^ ((i >> 4) & 7); }
MG_HD int swz2(int i) { return i ^ (((i >> 6) & 1) << 3); }
MG_HD int fft_slot(int m) { return (m & 1) * 8 + (m >> 1); }      // register slot of Z[l + 32 m]

// Twiddle tables (built once on the host in double precision, see fft_tables.h):
//   tw2[r * 8 + k]   = exp(-2 pi i r k / 64),   r = 0..7, k = 0..7
//   tw3[r * 64 + j]  = exp(-2 pi i r j / 512),  r = 0..7, j = 0..63
struct FftTables {
    float2 tw2[64];
    float2 tw3[512];
};

// ---- phases (lane l) --------------------------------------------------------------------
// pass 1: v[0..7] = inputs j + 64 r for j = l, v[8..15] for j = l + 32 (already loaded).
MG_HD void fft512_pass1_store(float2* v, float2* ex, int l) {
    dft8(v); dft8(v + 8);
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        ex[swz1(8 * l + r)] = v[r];
        ex[swz1(8 * (l + 32) + r)] = v[8 + r];
    }
}
MG_HD void fft512_pass2_load(float2* v, const float2* ex, int l) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        v[r] = ex[swz1(l + 64 * r)];
        v[8 + r] = ex[swz1(l + 32 + 64 * r)];
    }
}
MG_HD void fft512_pass2_store(float2* v, float2* ex, const FftTables& tb, int l) {
    const int k = l & 7;
#pragma unroll
    for (int r = 1; r < 8; ++r) {
        const float2 w = tb.tw2[r * 8 + k];
        v[r] = cmul(v[r], w);
        v[8 + r] = cmul(v[8 + r], w);
    }
    dft8(v); dft8(v + 8);
    const int a = l >> 3;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        ex[swz2(64 * a + 8 * r + k)] = v[r];
        ex[swz2(64 * (a + 4) + 8 * r + k)] = v[8 + r];
    }
}
MG_HD void fft512_pass3_load(float2* v, const float2* ex, int l) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        v[r] = ex[swz2(l + 64 * r)];
        v[8 + r] = ex[swz2(l + 32 + 64 * r)];
    }
}
MG_HD void fft512_pass3_finish(float2* v, const FftTables& tb, int l) {
#pragma unroll
    for (int r = 1; r < 8; ++r) {
        v[r] = cmul(v[r], tb.tw3[r * 64 + l]);
        v[8 + r] = cmul(v[8 + r], tb.tw3[r * 64 + l + 32]);
    }
    dft8(v); dft8(v + 8);
}

// ---- real <-> half-complex split -------------------------------------------------------------
// Forward (real FFT of 1024 samples packed as z[n] = x[2n] + i x[2n+1]):
//   X[k] = (Z[k] + conj Z[512-k]) / 2  +  W_1024^k * (-i/2) * (Z[k] - conj Z[512-k])
// `zk` = Z[k], `zp` = Z[(512-k) % 512], `w` = W_1024^k = (cos, -sin)(2 pi k / 1024).
// The common factor 1/2 is NOT applied here (callers fold it into the window).
MG_HD float2 rfft_split(float2 zk, float2 zp, float2 w) {
    const float sx = zk.x + zp.x, sy = zk.y - zp.y;      // Z[k] + conj Z[p]
    const float dx = zk.x - zp.x, dy = zk.y + zp.y;      // Z[k] - conj Z[p]
    return make_float2(sx + (dy * w.x + dx * w.y), sy + (dy * w.y - dx * w.x));
}
// Inverse: from the half-complex X (X[512] taken as `xp` when k == 0) build the packed spectrum
//   Z[k] = (X[k] + conj X[512-k]) + i * conj(W_1024^k) * (X[k] - conj X[512-k])
// (common factor 1/2 again left to the caller) so that ifft512(Z)[n] = x[2n] + i x[2n+1].
MG_HD float2 irfft_merge(float2 xk, float2 xp, float2 w) {
    const float sx = xk.x + xp.x, sy = xk.y - xp.y;      // X[k] + conj X[p]
    const float dx = xk.x - xp.x, dy = xk.y + xp.y;      // X[k] - conj X[p]
    // i * conj(w) * d,  conj(w) = (w.x, -w.y):  conj(w)*d = (dx w.x + dy w.y) + i (dy w.x - dx w.y)
    const float tx = dx * w.x + dy * w.y, ty = dy * w.x - dx * w.y;
    return make_float2(sx - ty, sy + tx);
}

}  // namespace mg
