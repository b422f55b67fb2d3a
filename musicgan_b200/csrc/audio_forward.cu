// Forward spectrogram transform for sm_100a:  waveform -> STFT -> Bark-weighted magnitude and
// instantaneous frequency, normalised per clip and cut into 512-frame chunks.
//
// Replaces reference music_gan/audio/functions.py:38-62 (wav_to_stft) and :65-94
// (stft_to_phase_magn, with :13-35 diff / unwrap / bark_magn_scale).
//
// Pipeline (all on the caller's stream; intermediate arrays are FRAME major [t][512] so every
// global access is a full 128-byte line; per-clip scratch is 4 KiB/frame and is meant to stay
// L2 resident between the stages):
//   k_stft<POLAR>        1 warp = 1 frame: reflect/mono staging in smem, Hann, 512-pt complex
//                        FFT (fft512.cuh) + real split, |X|*bark, atan2 -> phi, magn ; magn min/max
//   k_unwrap_aggregate   per 64-frame segment and bin: exact (fp64) sum of the unwrap adjustments
//   k_unwrap_carry       exclusive scan of the segment sums along time (per clip, per bin)
//   k_ifreq              u[t] = fl32(phi[t] + fl32(S[t])), IF[t] = fl32(u[t] - u[t-1]) ; IF min/max
//   k_normalise_chunk    (x - min) / (max - min) * 2 - 1, drop the head, transpose to [chunk][f][t]
//
// Numerics that are reproduced on purpose: float32 constants for pi,
// fmodf-with-sign-fix remainder, unwrap adjustments accumulated exactly (every adjustment is a
// float32 near +-2pi, so the fp64 sum is exact and therefore independent of summation order),
// one rounding to float32 per cumulative-sum output, no FMA contraction in those steps.
#include "common.cuh"
#include "fft512.cuh"
#include "fft_tables.h"
#include "tables.cuh"

#include <mutex>

namespace mg {

constexpr int kBins = 512;        // n_fft / 2 (Nyquist dropped, functions.py:62)
constexpr int kHop = 256;
constexpr int kNfft = 1024;
constexpr int kFramesPerCta = 16; // k_stft tile (16 frames: 69 KB of shared memory -> three CTAs per SM)
constexpr int kStftWarps = 8;
constexpr int kSeg = 64;          // frames per unwrap segment

__device__ DeviceTables g_tables;

// __device__ symbols exist once PER DEVICE: the tables are uploaded the first time each device is used (a process may
// drive several GPUs through the Python API's device= arguments)
constexpr int kMaxDevices = 64;
static std::once_flag g_tables_once[kMaxDevices];
static cudaError_t g_tables_status[kMaxDevices];
cudaError_t ensure_tables() {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= kMaxDevices) return cudaErrorInvalidDevice;
    std::call_once(g_tables_once[dev], [dev] {
        DeviceTables* h = new DeviceTables;
        build_fft_tables(&h->fft);
        build_split_twiddles(h->w1024);
        g_tables_status[dev] = cudaMemcpyToSymbol(g_tables, h, sizeof(DeviceTables));
        delete h;
    });
    return g_tables_status[dev];
}
const DeviceTables* device_tables_ptr() {
    void* p = nullptr;
    cudaGetSymbolAddress(&p, g_tables);
    return (const DeviceTables*)p;
}

// ------------------------------------------------------------------------------------------------
// k_stft
// ------------------------------------------------------------------------------------------------
struct StftSmem {
    float wav[kHop * (kFramesPerCta - 1) + kNfft];   // samples 256 t0 - 512 ... (reflect resolved)
    float2 win2[512];                                // (w[2n], w[2n+1]) * 0.5 / sqrt(sum w^2)
    FftTables fft;
    float2 w1024[512];
    float bark[kBins];
    float2 ex[kStftWarps][512];
    double red[kStftWarps];
    float redf[2][kStftWarps];
};

enum { STFT_POLAR = 0, STFT_COMPLEX = 1 };

template <int MODE>
__global__ void __launch_bounds__(kStftWarps * 32, 3)
k_stft(const float* __restrict__ wav, int64_t n_samples, int channels, int64_t clip_stride, int64_t n_frames,
       const float* __restrict__ window, const float* __restrict__ bark_gain, const DeviceTables* __restrict__ tables,
       float* __restrict__ out_a,      // POLAR: phi [clip][t][512]      COMPLEX: c64 [clip][t][512]
       float* __restrict__ out_b,      // POLAR: magn [clip][t][512]
       int* __restrict__ minmax_keys)  // POLAR: [clip][4] keys (0: magn min, 1: magn max)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    StftSmem& s = *reinterpret_cast<StftSmem*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int clip = blockIdx.y;
    const int64_t t0 = (int64_t)blockIdx.x * kFramesPerCta;
    const int frames_here = (int)min((int64_t)kFramesPerCta, n_frames - t0);

    // ---- window norm: sum w^2 (functions.py:53-59 normalized=True -> / sqrt(sum w^2)) ----
    double part = 0.0;
    for (int i = tid; i < kNfft; i += blockDim.x) { const double w = window[i]; part += w * w; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if (lane == 0) s.red[warp] = part;

    // ---- constant tables -> smem ----
    {
        const float2* src = reinterpret_cast<const float2*>(&tables->fft);
        float2* dst = reinterpret_cast<float2*>(&s.fft);
        for (int i = tid; i < (int)(sizeof(FftTables) / sizeof(float2)); i += blockDim.x) dst[i] = src[i];
        for (int i = tid; i < 512; i += blockDim.x) s.w1024[i] = tables->w1024[i];
        if (MODE == STFT_POLAR)
            for (int i = tid; i < kBins; i += blockDim.x) s.bark[i] = bark_gain[i];
    }

    // ---- stage the samples of this tile: mono mean (functions.py:49) + reflect pad (centre=True) ----
    {
        const float* base = wav + (int64_t)clip * clip_stride;
        const int64_t s0 = t0 * kHop - kNfft / 2;
        const int count = kHop * (frames_here - 1) + kNfft;
        const float inv_c = 1.0f;  // division below keeps torch.mean's rounding for any channel count
        (void)inv_c;
        for (int i = tid; i < count; i += blockDim.x) {
            int64_t p = s0 + i;
            if (p < 0) p = -p;
            if (p >= n_samples) p = 2 * (n_samples - 1) - p;
            float acc = base[p];
            for (int c = 1; c < channels; ++c) acc = __fadd_rn(acc, base[(int64_t)c * n_samples + p]);
            if (channels > 1) acc = __fdiv_rn(acc, (float)channels);
            s.wav[i] = acc;
        }
    }
    __syncthreads();
    {
        double tot = 0.0;
#pragma unroll
        for (int w = 0; w < kStftWarps; ++w) tot += s.red[w];
        // 0.5: the real-split leaves a factor 2 (fft512.cuh rfft_split)
        const float scale = (float)(0.5 / sqrt(tot));
        for (int i = tid; i < 512; i += blockDim.x)
            s.win2[i] = make_float2(window[2 * i] * scale, window[2 * i + 1] * scale);
    }
    __syncthreads();

    float mn = INFINITY, mx = -INFINITY;
    float2* ex = s.ex[warp];
    const int src_lane = (32 - lane) & 31;

    for (int fi = warp; fi < frames_here; fi += kStftWarps) {
        const int64_t t = t0 + fi;
        float2 v[16];
        {
            const float2* x2 = reinterpret_cast<const float2*>(s.wav + kHop * fi);
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const int n0 = lane + 64 * r, n1 = n0 + 32;
                const float2 a = x2[n0], wa = s.win2[n0], b = x2[n1], wb = s.win2[n1];
                v[r] = make_float2(a.x * wa.x, a.y * wa.y);
                v[8 + r] = make_float2(b.x * wb.x, b.y * wb.y);
            }
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

        const int64_t row = ((int64_t)clip * n_frames + t) * kBins;
#pragma unroll
        for (int m = 0; m < 16; ++m) {
            const int k = lane + 32 * m;
            const float2 offer = v[fft_slot(15 - m)];
            float px = __shfl_sync(0xffffffffu, offer.x, src_lane);
            float py = __shfl_sync(0xffffffffu, offer.y, src_lane);
            if (lane == 0) { const float2 own = v[fft_slot((16 - m) & 15)]; px = own.x; py = own.y; }
            const float2 X = rfft_split(v[fft_slot(m)], make_float2(px, py), s.w1024[k]);
            if (MODE == STFT_COMPLEX) {
                reinterpret_cast<float2*>(out_a)[row + k] = X;
            } else {
                // functions.py:69-72: abs, angle, bark gain
                const float mag = __fmul_rn(sqrtf(fmaf(X.x, X.x, X.y * X.y)), s.bark[k]);
                out_a[row + k] = atan2f(X.y, X.x);
                out_b[row + k] = mag;
                if (t >= 1) { mn = fminf(mn, mag); mx = fmaxf(mx, mag); }   // functions.py:77 drops frame 0
            }
        }
    }

    if (MODE == STFT_POLAR) {
        mn = warp_min(mn); mx = warp_max(mx);
        if (lane == 0) { s.redf[0][warp] = mn; s.redf[1][warp] = mx; }
        __syncthreads();
        if (tid == 0) {
#pragma unroll
            for (int w = 1; w < kStftWarps; ++w) { mn = fminf(mn, s.redf[0][w]); mx = fmaxf(mx, s.redf[1][w]); }
            if (mn <= mx) {
                atomicMin(&minmax_keys[clip * 4 + 0], float_key(mn));
                atomicMax(&minmax_keys[clip * 4 + 1], float_key(mx));
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// staged entry: polar form of a caller supplied complex STFT (functions.py:69-72)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_polar_from_stft(const float2* __restrict__ stft, int64_t n_frames, int64_t stride_f, int64_t stride_t,
                  int64_t batch_stride, const float* __restrict__ bark_gain,
                  float* __restrict__ phi, float* __restrict__ magn, int* __restrict__ minmax_keys) {
    __shared__ float redf[2][8];
    const int clip = blockIdx.y;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t total = n_frames * kBins;
    float mn = INFINITY, mx = -INFINITY;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + tid; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t t = i / kBins;
        const int f = (int)(i % kBins);
        const float2 X = stft[(int64_t)clip * batch_stride + f * stride_f + t * stride_t];
        const float mag = __fmul_rn(sqrtf(fmaf(X.x, X.x, X.y * X.y)), bark_gain[f]);
        phi[((int64_t)clip * n_frames + t) * kBins + f] = atan2f(X.y, X.x);
        magn[((int64_t)clip * n_frames + t) * kBins + f] = mag;
        if (t >= 1) { mn = fminf(mn, mag); mx = fmaxf(mx, mag); }
    }
    mn = warp_min(mn); mx = warp_max(mx);
    if (lane == 0) { redf[0][warp] = mn; redf[1][warp] = mx; }
    __syncthreads();
    if (tid == 0) {
        for (int w = 1; w < 8; ++w) { mn = fminf(mn, redf[0][w]); mx = fmaxf(mx, redf[1][w]); }
        if (mn <= mx) {
            atomicMin(&minmax_keys[clip * 4 + 0], float_key(mn));
            atomicMax(&minmax_keys[clip * 4 + 1], float_key(mx));
        }
    }
}

// ------------------------------------------------------------------------------------------------
// unwrap: segment sums, carry scan, instantaneous frequency
// ------------------------------------------------------------------------------------------------
__global__ void k_init_keys(int* __restrict__ keys, int batch) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < batch * 4) keys[i] = (i & 1) ? INT_MIN : INT_MAX;
}

void launch_init_keys(int* keys, int n_clips, cudaStream_t st) {
    k_init_keys<<<(n_clips * 4 + 127) / 128, 128, 0, st>>>(keys, n_clips);
}

// grid (n_seg, batch), block 512 (thread = bin)
__global__ void __launch_bounds__(kBins)
k_unwrap_aggregate(const float* __restrict__ phi, int64_t n_frames, int n_seg,
                   double* __restrict__ agg, float* __restrict__ bnd) {
    const int f = threadIdx.x, seg = blockIdx.x, clip = blockIdx.y;
    const int64_t tb = (int64_t)seg * kSeg, te = min(n_frames, tb + kSeg);
    const float* p = phi + (int64_t)clip * n_frames * kBins + f;
    float prev = 0.0f;
    if (tb > 0) prev = p[(tb - 1) * kBins];
    bnd[((int64_t)clip * n_seg + seg) * kBins + f] = prev;
    double S = 0.0;
    int64_t t = tb;
    if (t == 0) { prev = p[0]; t = 1; }        // diff() pads column 0 with zero (functions.py:14)
    for (; t + 8 <= te; t += 8) {
        float x[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = p[(t + i) * kBins];
#pragma unroll
        for (int i = 0; i < 8; ++i) { S += (double)unwrap_adjust(prev, x[i]); prev = x[i]; }
    }
    for (; t < te; ++t) { const float x = p[t * kBins]; S += (double)unwrap_adjust(prev, x); prev = x; }
    agg[((int64_t)clip * n_seg + seg) * kBins + f] = S;
}

// grid (batch), block 512: exclusive prefix over segments, in place
__global__ void __launch_bounds__(kBins)
k_unwrap_carry(double* __restrict__ agg, int n_seg) {
    const int f = threadIdx.x, clip = blockIdx.x;
    double* a = agg + (int64_t)clip * n_seg * kBins + f;
    double run = 0.0;
    int sg = 0;
    for (; sg + 4 <= n_seg; sg += 4) {
        double x[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) x[i] = a[(int64_t)(sg + i) * kBins];
#pragma unroll
        for (int i = 0; i < 4; ++i) { a[(int64_t)(sg + i) * kBins] = run; run += x[i]; }
    }
    for (; sg < n_seg; ++sg) { const double x = a[(int64_t)sg * kBins]; a[(int64_t)sg * kBins] = run; run += x; }
}

// grid (n_seg, batch), block 512.  Overwrites phi[t] (t >= 1) with IF[t] = u[t] - u[t-1]
// (functions.py:23 phi + cumsum, then :76 the time difference).
__global__ void __launch_bounds__(kBins)
k_ifreq(float* __restrict__ phi, int64_t n_frames, int n_seg,
        const double* __restrict__ carry, const float* __restrict__ bnd, int* __restrict__ minmax_keys) {
    __shared__ float redf[2][kBins / 32];
    const int f = threadIdx.x, seg = blockIdx.x, clip = blockIdx.y;
    const int lane = f & 31, warp = f >> 5;
    const int64_t tb = (int64_t)seg * kSeg, te = min(n_frames, tb + kSeg);
    float* p = phi + (int64_t)clip * n_frames * kBins + f;
    double S = carry[((int64_t)clip * n_seg + seg) * kBins + f];
    float prev, u_prev;
    int64_t t = tb;
    if (t == 0) { prev = p[0]; u_prev = __fadd_rn(prev, 0.0f); t = 1; }
    else { prev = bnd[((int64_t)clip * n_seg + seg) * kBins + f]; u_prev = __fadd_rn(prev, __double2float_rn(S)); }
    float mn = INFINITY, mx = -INFINITY;
    for (; t + 8 <= te; t += 8) {
        float x[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = p[(t + i) * kBins];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            S += (double)unwrap_adjust(prev, x[i]);
            const float u = __fadd_rn(x[i], __double2float_rn(S));
            const float d = __fsub_rn(u, u_prev);
            p[(t + i) * kBins] = d;
            mn = fminf(mn, d); mx = fmaxf(mx, d);
            u_prev = u; prev = x[i];
        }
    }
    for (; t < te; ++t) {
        const float x = p[t * kBins];
        S += (double)unwrap_adjust(prev, x);
        const float u = __fadd_rn(x, __double2float_rn(S));
        const float d = __fsub_rn(u, u_prev);
        p[t * kBins] = d;
        mn = fminf(mn, d); mx = fmaxf(mx, d);
        u_prev = u; prev = x;
    }
    mn = warp_min(mn); mx = warp_max(mx);
    if (lane == 0) { redf[0][warp] = mn; redf[1][warp] = mx; }
    __syncthreads();
    if (f == 0) {
        for (int w = 1; w < kBins / 32; ++w) { mn = fminf(mn, redf[0][w]); mx = fmaxf(mx, redf[1][w]); }
        if (mn <= mx) {
            atomicMin(&minmax_keys[clip * 4 + 2], float_key(mn));
            atomicMax(&minmax_keys[clip * 4 + 3], float_key(mx));
        }
    }
}

// ------------------------------------------------------------------------------------------------
// normalise + drop head + chunk (functions.py:84-92).  grid (n_cols/32, 512/128, batch), block 256
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float minmax_norm(float v, float mn, float mx) {
    // (x - min) / (max - min) * 2. - 1.   -- four separately rounded float32 ops
    return __fsub_rn(__fmul_rn(__fdiv_rn(__fsub_rn(v, mn), __fsub_rn(mx, mn)), 2.0f), 1.0f);
}

__global__ void __launch_bounds__(256)
k_normalise_chunk(const float* __restrict__ magn_ws, const float* __restrict__ if_ws, int64_t n_frames,
                  int64_t head, int64_t n_chunks, const int* __restrict__ minmax_keys,
                  float* __restrict__ magn, float* __restrict__ ifreq, float* __restrict__ minmax_out) {
    __shared__ float tile[2][32][129];
    const int tid = threadIdx.x;
    const int clip = blockIdx.z, f0 = blockIdx.y * 128;
    const int64_t c0 = (int64_t)blockIdx.x * 32;
    const float m_mn = key_float(minmax_keys[clip * 4 + 0]), m_mx = key_float(minmax_keys[clip * 4 + 1]);
    const float p_mn = key_float(minmax_keys[clip * 4 + 2]), p_mx = key_float(minmax_keys[clip * 4 + 3]);
    if (blockIdx.x == 0 && blockIdx.y == 0 && tid == 0 && minmax_out) {
        minmax_out[clip * 4 + 0] = m_mn; minmax_out[clip * 4 + 1] = m_mx;
        minmax_out[clip * 4 + 2] = p_mn; minmax_out[clip * 4 + 3] = p_mx;
    }
    const int64_t row0 = (int64_t)clip * n_frames + 1 + head + c0;   // frame of column c0
#pragma unroll 4
    for (int idx = tid; idx < 32 * 128; idx += 256) {
        const int i = idx >> 7, fb = idx & 127;
        const int64_t g = (row0 + i) * kBins + f0 + fb;
        tile[0][i][fb] = __ldcs(magn_ws + g);
        tile[1][i][fb] = __ldcs(if_ws + g);
    }
    __syncthreads();
    const int64_t chunk = c0 / kBins;
    const int col = (int)(c0 % kBins);
#pragma unroll 4
    for (int idx = tid; idx < 32 * 128; idx += 256) {
        const int fb = idx >> 5, i = idx & 31;
        const int64_t o = (((int64_t)clip * n_chunks + chunk) * kBins + f0 + fb) * kBins + col + i;
        __stcs(magn + o, minmax_norm(tile[0][i][fb], m_mn, m_mx));
        __stcs(ifreq + o, minmax_norm(tile[1][i][fb], p_mn, p_mx));
    }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
struct ForwardWs {
    float* phi; float* magn; double* agg; float* bnd; int* keys;
    size_t bytes;
};
static ForwardWs carve_forward_ws(void* ws, int64_t n_frames, int batch) {
    const int n_seg = (int)((n_frames + kSeg - 1) / kSeg);
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
    const size_t o_phi = take((size_t)batch * n_frames * kBins * sizeof(float));
    const size_t o_magn = take((size_t)batch * n_frames * kBins * sizeof(float));
    const size_t o_agg = take((size_t)batch * n_seg * kBins * sizeof(double));
    const size_t o_bnd = take((size_t)batch * n_seg * kBins * sizeof(float));
    const size_t o_keys = take((size_t)batch * 4 * sizeof(int));
    ForwardWs w;
    char* b = (char*)ws;
    w.phi = (float*)(b + o_phi); w.magn = (float*)(b + o_magn); w.agg = (double*)(b + o_agg);
    w.bnd = (float*)(b + o_bnd); w.keys = (int*)(b + o_keys); w.bytes = off;
    return w;
}

static int run_post_fft(const ForwardWs& w, int64_t n_frames, int batch, int64_t head, int64_t n_chunks,
                        float* magn, float* ifreq, float* minmax, cudaStream_t st) {
    const int n_seg = (int)((n_frames + kSeg - 1) / kSeg);
    { ProfScope ps("k_unwrap_aggregate", st);
      k_unwrap_aggregate<<<dim3(n_seg, batch), kBins, 0, st>>>(w.phi, n_frames, n_seg, w.agg, w.bnd); }
    { ProfScope ps("k_unwrap_carry", st);
      k_unwrap_carry<<<batch, kBins, 0, st>>>(w.agg, n_seg); }
    { ProfScope ps("k_ifreq", st);
      k_ifreq<<<dim3(n_seg, batch), kBins, 0, st>>>(w.phi, n_frames, n_seg, w.agg, w.bnd, w.keys); }
    if (n_chunks > 0) {
        ProfScope ps("k_normalise_chunk", st);
        k_normalise_chunk<<<dim3((unsigned)(n_chunks * kBins / 32), kBins / 128, batch), 256, 0, st>>>(
            w.magn, w.phi, n_frames, head, n_chunks, w.keys, magn, ifreq, minmax);
    }
    return check_launch("post_fft");
}

}  // namespace mg

using namespace mg;

extern "C" {

int mg_chunk_plan(int64_t n_samples, int hop, int nb_vec, int64_t* n_frames, int64_t* head, int64_t* n_chunks) {
    if (n_samples < 0 || hop <= 0 || nb_vec <= 0) return MG_ERR_BAD_ARG;
    const int64_t t = 1 + n_samples / hop;
    const int64_t cols = t - 1;
    const int64_t h = cols % nb_vec;
    if (n_frames) *n_frames = t;
    if (head) *head = h;
    if (n_chunks) *n_chunks = (cols - h) / nb_vec;
    return MG_OK;
}

size_t mg_stft_magif_workspace_bytes(int64_t n_samples, int batch) {
    if (n_samples <= 0 || batch <= 0) return 0;
    return carve_forward_ws(nullptr, 1 + n_samples / kHop, batch).bytes;
}
size_t mg_phase_magn_workspace_bytes(int64_t n_frames, int batch) {
    if (n_frames <= 0 || batch <= 0) return 0;
    return carve_forward_ws(nullptr, n_frames, batch).bytes;
}

static int launch_stft(int mode, const float* wav, int64_t n_samples, int channels, int64_t clip_stride, int batch,
                       const float* window, const float* bark, float* out_a, float* out_b, int* keys, cudaStream_t st) {
    cudaError_t e = ensure_tables();
    if (e != cudaSuccess) { set_last_cuda_error("tables", e); return MG_ERR_LAUNCH; }
    const int64_t n_frames = 1 + n_samples / kHop;
    const dim3 grid((unsigned)((n_frames + kFramesPerCta - 1) / kFramesPerCta), batch);
    const size_t smem = sizeof(StftSmem);
    ProfScope ps("k_stft", st);
    if (mode == STFT_POLAR) {
        cudaFuncSetAttribute(k_stft<STFT_POLAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);      // per device: set on every call
        k_stft<STFT_POLAR><<<grid, kStftWarps * 32, smem, st>>>(wav, n_samples, channels, clip_stride, n_frames, window, bark,
                                                                device_tables_ptr(), out_a, out_b, keys);
    } else {
        cudaFuncSetAttribute(k_stft<STFT_COMPLEX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        k_stft<STFT_COMPLEX><<<grid, kStftWarps * 32, smem, st>>>(wav, n_samples, channels, clip_stride, n_frames, window, bark,
                                                                  device_tables_ptr(), out_a, out_b, keys);
    }
    return check_launch("k_stft");
}

static int check_wav_args(const float* wav, int64_t n_samples, int channels, int64_t clip_stride, int batch, const float* window) {
    if (!wav || !window || batch <= 0 || channels <= 0) return MG_ERR_BAD_ARG;
    if (n_samples <= kNfft / 2) return MG_ERR_UNSUPPORTED;          // reflect pad needs N > n_fft/2 (torch raises too)
    if (clip_stride < (int64_t)channels * n_samples) return MG_ERR_BAD_ARG;
    if (batch > 65535) return MG_ERR_UNSUPPORTED;
    return MG_OK;
}

int mg_stft_magif_f32(const float* wav, int64_t n_samples, int channels, int64_t clip_stride, int batch,
                      const float* window, const float* bark_gain,
                      float* magn, float* ifreq, float* minmax,
                      void* ws, size_t ws_bytes, mgStream stream) {
    int rc = check_wav_args(wav, n_samples, channels, clip_stride, batch, window);
    if (rc) return rc;
    if (!bark_gain || !ws) return MG_ERR_BAD_ARG;
    int64_t n_frames, head, n_chunks;
    mg_chunk_plan(n_samples, kHop, kBins, &n_frames, &head, &n_chunks);
    if (n_chunks > 0 && (!magn || !ifreq)) return MG_ERR_BAD_ARG;
    if (((uintptr_t)ws & 255) != 0) return MG_ERR_BAD_ARG;
    ForwardWs w = carve_forward_ws(ws, n_frames, batch);
    if (ws_bytes < w.bytes) return MG_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    launch_init_keys(w.keys, batch, st);
    rc = launch_stft(STFT_POLAR, wav, n_samples, channels, clip_stride, batch, window, bark_gain, w.phi, w.magn, w.keys, st);
    if (rc) return rc;
    return run_post_fft(w, n_frames, batch, head, n_chunks, magn, ifreq, minmax, st);
}

int mg_stft_c64(const float* wav, int64_t n_samples, int channels, int64_t clip_stride, int batch,
                const float* window, float* out_c64, mgStream stream) {
    int rc = check_wav_args(wav, n_samples, channels, clip_stride, batch, window);
    if (rc) return rc;
    if (!out_c64) return MG_ERR_BAD_ARG;
    return launch_stft(STFT_COMPLEX, wav, n_samples, channels, clip_stride, batch, window, nullptr, out_c64, nullptr, nullptr,
                       (cudaStream_t)stream);
}

int mg_phase_magn_from_stft(const float* stft_c64, int64_t n_frames, int64_t stride_f, int64_t stride_t,
                            int64_t batch_stride, int batch, const float* bark_gain,
                            float* magn, float* ifreq, float* minmax,
                            void* ws, size_t ws_bytes, mgStream stream) {
    if (!stft_c64 || !bark_gain || !ws || batch <= 0 || n_frames < 2) return MG_ERR_BAD_ARG;
    if (batch > 65535) return MG_ERR_UNSUPPORTED;
    if (((uintptr_t)ws & 255) != 0 || ((uintptr_t)stft_c64 & 7) != 0) return MG_ERR_BAD_ARG;
    const int64_t cols = n_frames - 1, head = cols % kBins, n_chunks = (cols - head) / kBins;
    if (n_chunks > 0 && (!magn || !ifreq)) return MG_ERR_BAD_ARG;
    ForwardWs w = carve_forward_ws(ws, n_frames, batch);
    if (ws_bytes < w.bytes) return MG_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    launch_init_keys(w.keys, batch, st);
    const int64_t total = n_frames * kBins;
    const unsigned gx = (unsigned)min((int64_t)4096, (total + 255) / 256);
    { ProfScope ps("k_polar_from_stft", st);
      k_polar_from_stft<<<dim3(gx, batch), 256, 0, st>>>(reinterpret_cast<const float2*>(stft_c64), n_frames, stride_f, stride_t,
                                                         batch_stride, bark_gain, w.phi, w.magn, w.keys); }
    int rc = check_launch("k_polar_from_stft");
    if (rc) return rc;
    return run_post_fft(w, n_frames, batch, head, n_chunks, magn, ifreq, minmax, st);
}

}  // extern "C"
