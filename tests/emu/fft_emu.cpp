// Host emulation of the warp-level FFT phases in musicgan_b200/csrc/fft512.cuh:
// 32 lanes stepped in lock-step through the phases (the kernel separates them by __syncwarp).
// Checks (1) the 512-point FFT against a direct double DFT, (2) the real split / merge pair,
// (3) that every exchange access pattern is shared-memory bank-conflict free per half warp.
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../musicgan_b200/csrc/fft512.cuh"
#include "../../musicgan_b200/csrc/fft_tables.h"

using cd = std::complex<double>;
static int conflicts(int (*idx)(int lane, int r, int half), const char* name) {
    int bad = 0;
    for (int half = 0; half < 2; ++half)          // which butterfly (j = l or l + 32)
        for (int r = 0; r < 8; ++r)
            for (int hw = 0; hw < 2; ++hw) {      // half warp: 64-bit accesses go 16 lanes at a time
                int seen[16] = {0};
                for (int l = hw * 16; l < hw * 16 + 16; ++l) {
                    int b = idx(l, r, half) & 15;
                    if (seen[b]++) ++bad;
                }
            }
    std::printf("%-14s bank conflicts: %d\n", name, bad);
    return bad;
}
static int i_p1s(int l, int r, int h) { return mg::swz1(8 * (l + 32 * h) + r); }
static int i_p2l(int l, int r, int h) { return mg::swz1(l + 32 * h + 64 * r); }
static int i_p2s(int l, int r, int h) { return mg::swz2(64 * ((l >> 3) + 4 * h) + 8 * r + (l & 7)); }
static int i_p3l(int l, int r, int h) { return mg::swz2(l + 32 * h + 64 * r); }

int main() {
    mg::FftTables tb;
    mg::build_fft_tables(&tb);
    std::vector<float2> w1024(512);
    mg::build_split_twiddles(w1024.data());

    srand(1);
    std::vector<cd> z(512);
    for (auto& c : z) c = cd(rand() / (double)RAND_MAX - 0.5, rand() / (double)RAND_MAX - 0.5);

    float2 regs[32][16];
    std::vector<float2> ex(512);
    for (int l = 0; l < 32; ++l)
        for (int r = 0; r < 8; ++r) {
            regs[l][r] = make_float2((float)z[l + 64 * r].real(), (float)z[l + 64 * r].imag());
            regs[l][8 + r] = make_float2((float)z[l + 32 + 64 * r].real(), (float)z[l + 32 + 64 * r].imag());
        }
    for (int l = 0; l < 32; ++l) mg::fft512_pass1_store(regs[l], ex.data(), l);
    for (int l = 0; l < 32; ++l) mg::fft512_pass2_load(regs[l], ex.data(), l);
    for (int l = 0; l < 32; ++l) mg::fft512_pass2_store(regs[l], ex.data(), tb, l);
    for (int l = 0; l < 32; ++l) mg::fft512_pass3_load(regs[l], ex.data(), l);
    for (int l = 0; l < 32; ++l) mg::fft512_pass3_finish(regs[l], tb, l);

    double err = 0, ref = 0;
    std::vector<cd> Z(512);
    for (int k = 0; k < 512; ++k) {
        cd s = 0;
        for (int n = 0; n < 512; ++n) s += z[n] * std::polar(1.0, -2.0 * M_PI * n * k / 512.0);
        Z[k] = s;
    }
    for (int l = 0; l < 32; ++l)
        for (int m = 0; m < 16; ++m) {
            float2 g = regs[l][mg::fft_slot(m)];
            err = std::max(err, std::abs(cd(g.x, g.y) - Z[l + 32 * m]));
            ref = std::max(ref, std::abs(Z[l + 32 * m]));
        }
    std::printf("fft512 max abs err %.3e (max |Z| %.3e)\n", err, ref);
    int fail = err > 2e-5 * ref;

    // real split: x[2n] + i x[2n+1] = z[n]  ->  X = rfft(x)[0..511]
    double serr = 0, merr = 0;
    std::vector<cd> X(513);
    for (int k = 0; k <= 512; ++k) {
        cd s = 0;
        for (int n = 0; n < 512; ++n) {
            s += z[n].real() * std::polar(1.0, -2.0 * M_PI * (2 * n) * k / 1024.0);
            s += z[n].imag() * std::polar(1.0, -2.0 * M_PI * (2 * n + 1) * k / 1024.0);
        }
        X[k] = s;
    }
    for (int k = 0; k < 512; ++k) {
        int p = (512 - k) % 512;
        float2 zk = make_float2((float)Z[k].real(), (float)Z[k].imag());
        float2 zp = make_float2((float)Z[p].real(), (float)Z[p].imag());
        float2 x = mg::rfft_split(zk, zp, w1024[k]);
        serr = std::max(serr, std::abs(cd(0.5 * x.x, 0.5 * x.y) - X[k]));
        // merge back: needs X[k] and X[512-k]
        float2 xk = make_float2((float)X[k].real(), (float)X[k].imag());
        float2 xp = make_float2((float)X[512 - k].real(), (float)X[512 - k].imag());
        float2 zz = mg::irfft_merge(xk, xp, w1024[k]);
        merr = std::max(merr, std::abs(cd(0.5 * zz.x, 0.5 * zz.y) - Z[k]));
    }
    std::printf("rfft split max abs err %.3e, irfft merge max abs err %.3e\n", serr, merr);
    fail |= serr > 2e-5 * ref || merr > 2e-5 * ref;

    fail |= conflicts(i_p1s, "pass1 store");
    fail |= conflicts(i_p2l, "pass2 load");
    fail |= conflicts(i_p2s, "pass2 store");
    fail |= conflicts(i_p3l, "pass3 load");
    std::printf(fail ? "FAIL\n" : "OK\n");
    return fail;
}
