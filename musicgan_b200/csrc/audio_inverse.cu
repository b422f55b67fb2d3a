// Inverse spectrogram transform for sm_100a: (magnitude, instantaneous-frequency) images ->
// cumulative phase -> complex spectrum -> iSTFT (Hann 1024 / hop 256, centred) -> waveform.
//
// Replaces reference music_gan/audio/functions.py:97-137 (magn_phase_to_wav without the file write).
//
//   k_inv_phase_accumulate  phase affine map + STRICTLY SEQUENTIAL float32 running sum along time, one thread per
//                      (clip, bin) chain (:115-118; SURVEY B.3: a parallel or fp64 scan only reaches 44 dB on
//                      coherent phase) -- only the dependent FADD chain is serial; 128-byte row pieces are prefetched
//                      one block ahead so a chain never waits for memory
//   k_inv_polar        (m+1)/2 / bark (:111-112) and its per-clip min / max, % 2pi, magn * (cos, sin), [f][t] -> X[t][f]
//                      frame major, fully parallel (:120-123).  The division by (max - min) of :113 is a per-clip
//                      scalar and the iSTFT is linear, so it is applied to the output samples in k_istft: no separate
//                      min / max pass over the magnitude planes
//   k_istft            per frame: half-complex -> packed 512-pt spectrum, inverse FFT (fft512.cuh),
//                      Hann, overlap-add of 4 frames in registers (ascending frame order), divide by the
//                      window envelope (and the magnitude range), trim n_fft/2 at both ends   (:125-137)
#include "common.cuh"
#include "fft512.cuh"
#include "fft_tables.h"
#include "tables.cuh"

namespace mg {

constexpr int kIBins = 512;
constexpr int kIHop = 256;
constexpr int kINfft = 1024;

__device__ __forceinline__ int64_t img_index(int clip, int imgs, int W, int ch, int f, int64_t tt) {
    const int i = (int)(tt / W), w = (int)(tt % W);
    return ((((int64_t)clip * imgs + i) * 2 + ch) * kIBins + f) * (int64_t)W + w;
}

// functions.py:111-112
__device__ __forceinline__ float magn_unscaled(float m, float gain) {
    return __fdiv_rn(__fdiv_rn(__fadd_rn(m, 1.0f), 2.0f), gain);
}

// :115-118 -- the phase affine map and the STRICTLY SEQUENTIAL float32 running sum along time, one thread per
// (clip, bin) chain.  Only the cheap dependent chain lives here (one FADD per step); everything per element (fmodf,
// sincosf, magnitude) is done in parallel by k_inv_polar.  acc [n_clips][imgs][512][W] fp32 (same layout as the input).
// A thread walks its row in blocks of 32 floats (one 128-byte line) and loads block k+1 before it sums block k: the
// chain then costs its 4-cycle adds, not one memory round trip per block.
// grid (512 / kAccThreads, n_clips), block kAccThreads
constexpr int kAccThreads = 32;

__device__ __forceinline__ float phase_affine(float v) {       // :115  ((p + 1) / 2 * 2) * pi - pi, float32 constants
    return __fsub_rn(__fmul_rn(__fmul_rn(__fdiv_rn(__fadd_rn(v, 1.0f), 2.0f), 2.0f), kPiF), kPiF);
}

__global__ void __launch_bounds__(kAccThreads)
k_inv_phase_accumulate(const float* __restrict__ mp, int imgs, int W, float* __restrict__ acc_out) {
    const int clip = blockIdx.y, f = blockIdx.x * kAccThreads + threadIdx.x;
    float acc = 0.0f;
    bool first = true;
    const bool vec = (W & 31) == 0;
    for (int i = 0; i < imgs; ++i) {
        const float* src = mp + ((((int64_t)clip * imgs + i) * 2 + 1) * kIBins + f) * (int64_t)W;
        float* dst = acc_out + (((int64_t)clip * imgs + i) * kIBins + f) * (int64_t)W;
        int w = 0;
        if (vec) {
            float4 cur[8], nxt[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) cur[j] = __ldcs(reinterpret_cast<const float4*>(src) + j);
            for (; w < W; w += 32) {
                if (w + 32 < W) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) nxt[j] = __ldcs(reinterpret_cast<const float4*>(src + w + 32) + j);
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float v[4] = {cur[j].x, cur[j].y, cur[j].z, cur[j].w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float p = phase_affine(v[e]);
                        acc = first ? p : __fadd_rn(acc, p);
                        first = false;
                        v[e] = acc;
                    }
                    __stcs(reinterpret_cast<float4*>(dst + w) + j, make_float4(v[0], v[1], v[2], v[3]));
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) cur[j] = nxt[j];
            }
        }
        for (; w < W; ++w) {
            const float p = phase_affine(src[w]);
            acc = first ? p : __fadd_rn(acc, p);
            first = false;
            dst[w] = acc;
        }
    }
}

// :111-112,120-123 -- per element: magnitude de-normalisation (without the per-clip range, see k_istft) and its min / max,
// phase % 2pi, magn * (cos, sin); [f][t] tiles are transposed through shared memory so that X[t][f] (frame major, what
// k_istft reads) is written in full lines.
// grid (ceil(Wt / 32), 16, n_clips), block 256
__global__ void __launch_bounds__(256)
k_inv_polar(const float* __restrict__ mp, const float* __restrict__ acc_in, int imgs, int W, const float* __restrict__ bark,
            int* __restrict__ keys, float2* __restrict__ X) {
    __shared__ float tp[32][33];
    __shared__ float tm[32][33];
    __shared__ float redf[2][8];
    const int clip = blockIdx.z, f0 = blockIdx.y * 32;
    const int64_t Wt = (int64_t)imgs * W, tt0 = (int64_t)blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    float mn = INFINITY, mx = -INFINITY;
    for (int r = ty; r < 32; r += 8) {              // r = bin within the tile, tx = time within the tile
        const int64_t tt = tt0 + tx;
        float a = 0.0f, m = 0.0f;
        if (tt < Wt) {
            const int i = (int)(tt / W), w = (int)(tt % W);
            a = __ldcs(acc_in + (((int64_t)clip * imgs + i) * kIBins + f0 + r) * (int64_t)W + w);
            m = __ldcs(mp + img_index(clip, imgs, W, 0, f0 + r, tt));
        }
        tp[r][tx] = a; tm[r][tx] = m;
    }
    __syncthreads();
    const float gain = bark[f0 + tx];
    for (int c = ty; c < 32; c += 8) {              // c = time within the tile, tx = bin within the tile
        const int64_t tt = tt0 + c;
        if (tt < Wt) {
            const float ph = remainder_pos(tp[tx][c], kTwoPiF);                   // :120
            float sn, cs;
            sincosf(ph, &sn, &cs);
            const float m = magn_unscaled(tm[tx][c], gain);
            mn = fminf(mn, m); mx = fmaxf(mx, m);
            X[((int64_t)clip * Wt + tt) * kIBins + f0 + tx] = make_float2(__fmul_rn(m, cs), __fmul_rn(m, sn));
        }
    }
    mn = warp_min(mn); mx = warp_max(mx);
    if (tx == 0) { redf[0][ty] = mn; redf[1][ty] = mx; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) { mn = fminf(mn, redf[0][w]); mx = fmaxf(mx, redf[1][w]); }
        if (mn <= mx) {
            atomicMin(&keys[clip * 4 + 0], float_key(mn));
            atomicMax(&keys[clip * 4 + 1], float_key(mx));
        }
    }
}

// ------------------------------------------------------------------------------------------------
constexpr int kIstftWarps = 8;
constexpr int kHopsPerWarp = 16;

struct IstftSmem {
    float2 win2[512];        // (w[2n], w[2n+1]) * sqrt(sum w^2) / 1024
    float wsq[kINfft];       // w^2 (envelope terms)
    FftTables fft;
    float2 w1024[512];
    float2 ex[kIstftWarps][512];
    double red[kIstftWarps];
};

// grid (ceil(n_hops / (8*16)), n_clips), block 256
__global__ void __launch_bounds__(kIstftWarps * 32, 3)
k_istft(const float2* __restrict__ X, int64_t Wt, const float* __restrict__ window,
        const DeviceTables* __restrict__ tables, const int* __restrict__ keys, float* __restrict__ wav) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    IstftSmem& s = *reinterpret_cast<IstftSmem*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int clip = blockIdx.y;
    const int64_t n_hops = Wt - 1;

    double part = 0.0;
    for (int i = tid; i < kINfft; i += blockDim.x) { const double w = window[i]; part += w * w; s.wsq[i] = __fmul_rn(window[i], window[i]); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if (lane == 0) s.red[warp] = part;
    {
        const float2* src = reinterpret_cast<const float2*>(&tables->fft);
        float2* dst = reinterpret_cast<float2*>(&s.fft);
        for (int i = tid; i < (int)(sizeof(FftTables) / sizeof(float2)); i += blockDim.x) dst[i] = src[i];
        for (int i = tid; i < 512; i += blockDim.x) s.w1024[i] = tables->w1024[i];
    }
    __syncthreads();
    {
        double tot = 0.0;
#pragma unroll
        for (int w = 0; w < kIstftWarps; ++w) tot += s.red[w];
        // * sqrt(sum w^2): inverse_spectrogram(normalized=True); / 1024: irfft norm (the merge keeps a factor 2
        // and the forward-FFT-with-conjugates trick a factor 512)
        const float scale = (float)(sqrt(tot) / 1024.0);
        for (int i = tid; i < 512; i += blockDim.x)
            s.win2[i] = make_float2(window[2 * i] * scale, window[2 * i + 1] * scale);
    }
    __syncthreads();

    const int64_t h0 = ((int64_t)blockIdx.x * kIstftWarps + warp) * kHopsPerWarp;
    if (h0 >= n_hops) return;
    const int64_t h1 = min(n_hops, h0 + kHopsPerWarp);
    float2* ex = s.ex[warp];
    const float2* Xc = X + (int64_t)clip * Wt * kIBins;
    float* out = wav + (int64_t)clip * n_hops * kIHop;
    // :113 magn / (max - min): a per-clip scalar, applied to the output samples (the transform is linear)
    const float range = __fsub_rn(key_float(keys[clip * 4 + 1]), key_float(keys[clip * 4 + 0]));

    float2 acc[16];
#pragma unroll
    for (int m = 0; m < 16; ++m) acc[m] = make_float2(0.f, 0.f);

    for (int64_t t = h0 - 1; t <= h1 + 1; ++t) {
        // slide the 1024-sample accumulation window forward by one hop (= 4 slots of 64 samples)
#pragma unroll
        for (int m = 0; m < 12; ++m) acc[m] = acc[m + 4];
#pragma unroll
        for (int m = 12; m < 16; ++m) acc[m] = make_float2(0.f, 0.f);

        if (t >= 0 && t < Wt) {
            const float2* row = Xc + t * kIBins;
            float2 v[16];
#pragma unroll
            for (int m = 0; m < 16; ++m) {
                const int k = lane + 32 * m;
                float2 xk = row[k];
                float2 xp = (k == 0) ? make_float2(0.f, 0.f) : row[kIBins - k];   // Nyquist row is zero (:125-126)
                if (k == 0) xk.y = 0.f;                                          // C2R ignores Im X[0]
                const float2 z = irfft_merge(xk, xp, s.w1024[k]);
                v[fft_slot(m)] = make_float2(z.x, -z.y);                         // conj: inverse via forward FFT
            }
            fft512_pass1_store(v, ex, lane);
            __syncwarp();
            fft512_pass2_load(v, ex, lane);
            __syncwarp();
            fft512_pass2_store(v, ex, s.fft, lane);
            __syncwarp();
            fft512_pass3_load(v, ex, lane);
            __syncwarp();
            fft512_pass3_finish(v, s.fft, lane);
#pragma unroll
            for (int m = 0; m < 16; ++m) {
                const float2 z = v[fft_slot(m)];
                const float2 w = s.win2[lane + 32 * m];
                // x[2n] = Re, x[2n+1] = -Im (conj back); frames are added in ascending t
                acc[m].x = __fadd_rn(acc[m].x, z.x * w.x);
                acc[m].y = __fadd_rn(acc[m].y, -z.y * w.y);
            }
        }
        const int64_t h = t - 2;
        if (h >= h0 && h < h1) {
#pragma unroll
            for (int m = 0; m < 4; ++m) {
                const int q0 = 2 * lane + 64 * m;
                float e0 = 0.f, e1 = 0.f;
#pragma unroll
                for (int i = 3; i >= 0; --i) {           // ascending frame index t' = h + 2 - i
                    const int64_t tf = h + 2 - i;
                    if (tf >= 0 && tf < Wt) { e0 = __fadd_rn(e0, s.wsq[q0 + 256 * i]); e1 = __fadd_rn(e1, s.wsq[q0 + 1 + 256 * i]); }
                }
                float2 o = make_float2(__fdiv_rn(acc[m].x, __fmul_rn(e0, range)), __fdiv_rn(acc[m].y, __fmul_rn(e1, range)));
                __stcs(reinterpret_cast<float2*>(out + h * kIHop + q0), o);
            }
        }
    }
}

struct InverseWs { float2* X; float* acc; int* keys; size_t bytes; };
static InverseWs carve_inverse_ws(void* ws, int n_clips, int64_t Wt) {
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
    const size_t o_x = take((size_t)n_clips * Wt * kIBins * sizeof(float2));
    const size_t o_a = take((size_t)n_clips * Wt * kIBins * sizeof(float));
    const size_t o_k = take((size_t)n_clips * 4 * sizeof(int));
    InverseWs w; char* b = (char*)ws;
    w.X = (float2*)(b + o_x); w.acc = (float*)(b + o_a); w.keys = (int*)(b + o_k); w.bytes = off;
    return w;
}

}  // namespace mg

using namespace mg;

extern "C" {

size_t mg_istft_workspace_bytes(int n_clips, int imgs_per_clip, int width) {
    if (n_clips <= 0 || imgs_per_clip <= 0 || width <= 0) return 0;
    return carve_inverse_ws(nullptr, n_clips, (int64_t)imgs_per_clip * width).bytes;
}

int mg_istft_from_magif_f32(const float* magn_phase, int n_clips, int imgs_per_clip, int width,
                            const float* window, const float* bark_gain,
                            float* wav, void* ws, size_t ws_bytes, mgStream stream) {
    if (!magn_phase || !window || !bark_gain || !wav || !ws) return MG_ERR_BAD_ARG;
    if (n_clips <= 0 || imgs_per_clip <= 0 || width <= 0) return MG_ERR_BAD_ARG;
    const int64_t Wt = (int64_t)imgs_per_clip * width;
    if (Wt < 2) return MG_ERR_UNSUPPORTED;
    if (n_clips > 65535) return MG_ERR_UNSUPPORTED;
    if (((uintptr_t)ws & 255) != 0 || ((uintptr_t)wav & 7) != 0) return MG_ERR_BAD_ARG;
    InverseWs w = carve_inverse_ws(ws, n_clips, Wt);
    if (ws_bytes < w.bytes) return MG_ERR_WORKSPACE;
    cudaError_t e = ensure_tables();
    if (e != cudaSuccess) { set_last_cuda_error("tables", e); return MG_ERR_LAUNCH; }
    cudaStream_t st = (cudaStream_t)stream;
    launch_init_keys(w.keys, n_clips, st);
    { ProfScope ps("k_inv_phase_accumulate", st);
      k_inv_phase_accumulate<<<dim3(kIBins / kAccThreads, n_clips), kAccThreads, 0, st>>>(magn_phase, imgs_per_clip, width, w.acc); }
    { ProfScope ps("k_inv_polar", st);
      k_inv_polar<<<dim3((unsigned)((Wt + 31) / 32), 16, n_clips), 256, 0, st>>>(magn_phase, w.acc, imgs_per_clip, width, bark_gain, w.keys, w.X); }
    const int64_t n_hops = Wt - 1;
    const unsigned gh = (unsigned)((n_hops + kIstftWarps * kHopsPerWarp - 1) / (kIstftWarps * kHopsPerWarp));
    cudaFuncSetAttribute(k_istft, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(IstftSmem));      // per device
    { ProfScope ps("k_istft", st);
      k_istft<<<dim3(gh, n_clips), kIstftWarps * 32, sizeof(IstftSmem), st>>>(w.X, Wt, window, device_tables_ptr(), w.keys, wav); }
    return check_launch("istft");
}

}  // extern "C"
