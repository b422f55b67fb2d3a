// 3x3 / stride 1 / pad 1 convolutions of the ProGAN blocks as tcgen05 implicit GEMMs (sm_100a).
//
// Replaces the nn.Conv2d(3x3, s1, p1) calls of reference music_gan/networks/generator.py:16-22,31-37
// and discriminator.py:15-21,26-32 (cuDNN in the reference) -- forward (fprop), data gradient (dgrad =
// the same kernel on flipped / transposed weights) and weight gradient (wgrad, conv_wgrad.cu).
//
// Layout: activations NHWC bf16 (torch channels_last), fp32 accumulation in TMEM, fp32 master weights
// packed to bf16 on the fly.
//
// One CTA works on output tiles of 16 x 8 pixels of one image (GEMM M = 128):
//   * producer warps stage the 18 x 10 input halo of the tile ONCE into shared memory in the canonical
//     no-swizzle UMMA layout [channel chunk][halo position][8 ch] (umma.cuh).  Zero padding, image borders
//     and the optional nearest 2x upsampling of the input (generator.py:26-29 folded into the read) are
//     resolved here, so the upsampled tensor never exists in HBM;
//   * the nine filter taps are nine START ADDRESSES into that halo tile: tap (ky,kx) starts at halo position
//     ky*10+kx, the 16 tile rows are the sixteen 8-row core-matrix groups, stride byte offset = one halo
//     row (160 B).  No im2col buffer, every input byte is fetched once per tile;
//   * weights of the CTA's output-channel slice stay resident in shared memory for the whole launch;
//   * one thread issues 9 * Cin/16 tcgen05.mma (M128 x N=slice x K16) per tile into one of two TMEM
//     accumulators; tcgen05.commit releases the halo slot and hands the accumulator to the epilogue warps;
//   * epilogue warps: tcgen05.ld -> + bias -> LeakyReLU(0.2) -> PixelNorm (per-pixel channel RMS, in fp32,
//     layers.py:11-17) -> bf16 -> NHWC store.  The thread that owns TMEM lane m owns pixel m of the tile and
//     sees all its channels, so PixelNorm needs no cross-thread traffic.
#include "common.cuh"
#include "umma.cuh"
#include "conv_common.cuh"
#include <cstdlib>

namespace mg {
using namespace umma;

struct ConvParams {
    const __nv_bfloat16* x;      // [B][Hin][Win][Cin]
    const uint4* wpack;          // packed bf16 weights (pack_weights)
    const float* bias;           // [Cout] or null
    __nv_bfloat16* y;            // [B][H][W][Cout]
    float* inv_norm;             // [B][H][W] 1/sqrt(mean_c(v^2)+eps) (pixelnorm, optional)
    int B, H, W, Hin, Win, Cin, Cout;
    int upsample, lrelu, pixelnorm;
    int tiles_x, tiles_y, n_tiles;
    int Nt, stages, tmem_cols, tiles_per_img;
    ItemDiv idiv;
    FastDiv div_img, div_tx;
    int consumer_fence;
    int epi_warps;               // 4 or 8 epilogue warps; producers are the next 4 warps, then the MMA warp
};

// warp roles: [0, E) epilogue (E = 4, or 8 = two per TMEM lane quarter with half the columns each), [E, E+4) producers,
// warp E+4 MMA issue + TMEM alloc.  Block size (E + 5) * 32.
constexpr int kConvThreads = 13 * 32;                // upper bound (E = 8)

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

__global__ void __launch_bounds__(kConvThreads, 2)
k_conv3x3(const ConvParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nch = p.Cin >> 3;
    const int slice = blockIdx.y;
    const int n0 = slice * p.Nt;
    const int nt = min(p.Nt, p.Cout - n0);
    const int kEpiWarps = p.epi_warps, kProdWarp0 = kEpiWarps, kMmaWarp = kEpiWarps + 4;

    uint4* sW = reinterpret_cast<uint4*>(smem);
    uint4* sA0 = sW + 9 * nch * nt;
    float* sBias = reinterpret_cast<float*>(sA0 + (size_t)p.stages * nch * kHaloPitch);      // 16-byte aligned
    float* sPart = sBias + ((nt + 3) & ~3);                                                  // [2][128] PixelNorm partial sums
    uint64_t* bars = reinterpret_cast<uint64_t*>(sPart + 256);
    uint64_t* full_a = bars;                       // [kMaxStages]
    uint64_t* empty_a = bars + kMaxStages;         // [kMaxStages]
    uint64_t* tmem_full = bars + 2 * kMaxStages;   // [2]
    uint64_t* tmem_empty = tmem_full + 2;          // [2]
    uint64_t* w_full = tmem_full + 4;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 5);

    if (tid == 0) {
        for (int i = 0; i < kMaxStages; ++i) { mbar_init(&full_a[i], 128); mbar_init(&empty_a[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], kEpiWarps * 32); }
        mbar_init(w_full, 128);
        mbar_fence_init();
    }
    if (warp == kMmaWarp) tmem_alloc(tmem_slot, p.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int acc_stride = p.tmem_cols >> 1;

    if (warp >= kProdWarp0 && warp < kMmaWarp) {
        // ================= producers =================
        const int pt = tid - kProdWarp0 * 32;
        {   // resident weights of this slice + bias
            const uint4* src = p.wpack + (size_t)9 * nch * n0;
            const int total = 9 * nch * nt;
            // all chunks in flight at once (the deep layers carry up to 140 KB of weights per slice: a synchronous
            // load loop here used to dominate the run time of the small-spatial layers)
            const uint32_t sw_addr = smem_u32(sW);
            for (int i = pt; i < total; i += 128) cp_async16(sw_addr + (uint32_t)i * 16u, src + i, 16u);
            for (int i = pt; i < nt; i += 128) sBias[i] = p.bias ? p.bias[n0 + i] : 0.0f;
            cp_async_arrive(w_full);
        }
        const int items = kHaloPos * nch;
        // Per-thread item table, built once: for each 16-byte chunk this thread copies per tile, where it comes from
        // relative to the tile origin (tile independent, because tiles start at even rows / columns), where it goes in
        // the halo slot, and its halo coordinates (for the image-border test).  Up to kItemRegs items live in registers;
        // wider layers (Cin > 64) take the generic path for the remaining items.
        constexpr int kItemRegs = 12;
        int rel[kItemRegs];            // source offset in uint4 units from the tile origin
        uint32_t dsto[kItemRegs];      // (hy << 24) | (hx << 16) | smem chunk index
        const int sy_step = p.upsample ? kTileH / 2 : kTileH, sx_step = p.upsample ? kTileW / 2 : kTileW;
#pragma unroll
        for (int k = 0; k < kItemRegs; ++k) {
            const int i = pt + k * 128;
            rel[k] = 0; dsto[k] = 0xFFFFFFFFu;
            if (i < items) {
                const int pos = (int)(((unsigned)i * p.idiv.magic) >> 20), c = i - pos * nch;
                const int hy = (pos * 6554) >> 16, hx = pos - hy * kHaloW;
                const int ry = p.upsample ? ((hy - 1) >> 1) : hy - 1, rx = p.upsample ? ((hx - 1) >> 1) : hx - 1;
                rel[k] = (ry * p.Win + rx) * nch + c;
                dsto[k] = ((uint32_t)hy << 24) | ((uint32_t)hx << 16) | (uint32_t)(c * kHaloPitch + pos);
            }
        }
        int slot = 0; uint32_t ph = 0;
        for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
            mbar_wait(&empty_a[slot], ph ^ 1u);
            const int b = fast_div(tile, p.div_img);
            const int tr = tile - b * p.tiles_per_img;
            const int tyi = fast_div(tr, p.div_tx), txi = tr - tyi * p.tiles_x;
            const int ty0 = tyi * kTileH - 1, tx0 = txi * kTileW - 1;
            const uint32_t dst = smem_u32(sA0 + (size_t)slot * nch * kHaloPitch);
            const uint4* img = reinterpret_cast<const uint4*>(p.x + (size_t)b * p.Hin * p.Win * p.Cin);
            const uint4* org = img + ((size_t)(tyi * sy_step) * p.Win + txi * sx_step) * nch;
            // asynchronous 16-byte copies (zero fill outside the image); nothing is waited for here, so the loads of
            // up to `stages` tiles are in flight per CTA
#pragma unroll
            for (int k = 0; k < kItemRegs; ++k) {
                if (dsto[k] != 0xFFFFFFFFu) {
                    const int iy = ty0 + (int)(dsto[k] >> 24), ix = tx0 + (int)((dsto[k] >> 16) & 0xFF);
                    const bool ok = (unsigned)iy < (unsigned)p.H && (unsigned)ix < (unsigned)p.W;
                    cp_async16(dst + (dsto[k] & 0xFFFFu) * 16u, ok ? (const void*)(org + rel[k]) : (const void*)img, ok ? 16u : 0u);
                }
            }
            for (int i = pt + kItemRegs * 128; i < items; i += 128) {
                const int pos = (int)(((unsigned)i * p.idiv.magic) >> 20), c = i - pos * nch;
                const int hy = (pos * 6554) >> 16, hx = pos - hy * kHaloW;
                const int iy = ty0 + hy, ix = tx0 + hx;
                const bool ok = (unsigned)iy < (unsigned)p.H && (unsigned)ix < (unsigned)p.W;
                const int sy = p.upsample ? (iy >> 1) : iy, sx = p.upsample ? (ix >> 1) : ix;
                cp_async16(dst + (uint32_t)(c * kHaloPitch + pos) * 16u,
                           ok ? (const void*)(img + ((size_t)sy * p.Win + sx) * nch + c) : (const void*)img, ok ? 16u : 0u);
            }
            cp_async_arrive(&full_a[slot]);
            if (++slot == p.stages) { slot = 0; ph ^= 1u; }
        }
    } else if (warp == kMmaWarp) {
        // ================= MMA issue (whole warp walks the loop, lane 0 issues) =================
        // Descriptors only differ in their 14-bit start-address field, so they are formed by integer additions:
        //   A: slot base + tap offset (ky*10 + kx positions) + k-step * 2 * kHaloPitch      (16-byte units)
        //   B: weights are packed [tap][k chunk][n] -> the descriptor simply advances by 2*nt per MMA
        const uint32_t idesc = instr_desc_bf16(nt, false, false);
        const uint64_t a_desc0 = smem_desc(smem_u32(sA0), kHaloPitch * 16u, kHaloW * 16u);
        const uint64_t b_desc0 = smem_desc(smem_u32(sW), (uint32_t)nt * 16u, 128u);
        const uint32_t slot_units = (uint32_t)(nch * kHaloPitch), b_step = (uint32_t)(2 * nt);
        const int ksteps = nch >> 1;
        mbar_wait(w_full, 0);
        int slot = 0, acc = 0; uint32_t ph = 0, aph = 0;
        for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
            mbar_wait(&tmem_empty[acc], aph ^ 1u);
            mbar_wait(&full_a[slot], ph);
            if (p.consumer_fence) fence_proxy_async();      // (debug switch) cp.async-written operands -> async proxy
            tc_fence_after();
            if (lane == 0) {
                const uint32_t d = tmem_base + acc * acc_stride;
                const uint64_t da_slot = a_desc0 + (uint64_t)(slot * slot_units);
                uint64_t db = b_desc0;
                uint32_t accum = 0;
#pragma unroll
                for (int tap = 0; tap < 9; ++tap) {
                    uint64_t da = da_slot + (uint64_t)((tap / 3) * kHaloW + (tap % 3));
                    for (int kk = 0; kk < ksteps; ++kk) {
                        mma_bf16(d, da, db, idesc, accum);
                        accum = 1;
                        da += 2 * kHaloPitch;
                        db += b_step;
                    }
                }
                mma_commit(&empty_a[slot]);
                mma_commit(&tmem_full[acc]);
            }
            __syncwarp();
            if (++slot == p.stages) { slot = 0; ph ^= 1u; }
            if (++acc == 2) { acc = 0; aph ^= 1u; }
        }
    } else {
        // ================= epilogue: 8 warps; warp = (half << 2) | quarter =================
        const int quarter = warp & 3, half = warp >> 2;
        const int m = quarter * 32 + lane;           // TMEM lane == pixel of the tile
        const int ry = m >> 3, rx = m & 7;
        // columns of this half, in units of 16: the first half takes the larger share
        const int units = nt >> 4;
        const bool split = kEpiWarps == 8;
        const int u0 = (!split || half == 0) ? 0 : (units + 1) >> 1, u1 = !split ? units : (half == 0 ? (units + 1) >> 1 : units);
        const float inv_c = 1.0f / (float)nt;
        int acc = 0; uint32_t aph = 0;
        for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
            const int b = fast_div(tile, p.div_img);
            const int tr = tile - b * p.tiles_per_img;
            const int tyi = fast_div(tr, p.div_tx), txi = tr - tyi * p.tiles_x;
            const int oy = tyi * kTileH + ry, ox = txi * kTileW + rx;
            const bool valid = oy < p.H && ox < p.W;
            mbar_wait(&tmem_full[acc], aph);
            tc_fence_after();
            const uint32_t taddr = tmem_base + acc * acc_stride + ((uint32_t)(quarter * 32) << 16);
            float scale = 1.0f;
            if (p.pixelnorm) {
                float ss = 0.0f;
                for (int u = u0; u < u1; ++u) {
                    float v[16];
                    tmem_ld16(taddr + u * 16, v);
                    tmem_wait_ld();
                    const float4* b4 = reinterpret_cast<const float4*>(sBias + u * 16);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float4 bb = b4[j];
                        float t0 = v[4 * j] + bb.x, t1 = v[4 * j + 1] + bb.y, t2 = v[4 * j + 2] + bb.z, t3 = v[4 * j + 3] + bb.w;
                        if (p.lrelu) { t0 = fmaxf(t0, 0.2f * t0); t1 = fmaxf(t1, 0.2f * t1); t2 = fmaxf(t2, 0.2f * t2); t3 = fmaxf(t3, 0.2f * t3); }
                        ss = fmaf(t0, t0, ss); ss = fmaf(t1, t1, ss); ss = fmaf(t2, t2, ss); ss = fmaf(t3, t3, ss);
                    }
                }
                float tot = ss;
                if (split) {
                    sPart[half * 128 + m] = ss;
                    named_bar_sync(1 + quarter, 64);              // the two warps that share this TMEM lane quarter
                    tot = sPart[m] + sPart[128 + m];
                    named_bar_sync(1 + quarter, 64);              // sPart may be overwritten by the next tile
                }
                scale = 1.0f / sqrtf(tot * inv_c + 1e-8f);
                if (valid && half == 0 && p.inv_norm) p.inv_norm[((size_t)b * p.H + oy) * p.W + ox] = scale;
            }
            __nv_bfloat16* dst = p.y + (((size_t)b * p.H + oy) * p.W + ox) * p.Cout + n0;
            for (int u = u0; u < u1; ++u) {
                float v[16];
                tmem_ld16(taddr + u * 16, v);
                tmem_wait_ld();
                const float4* b4 = reinterpret_cast<const float4*>(sBias + u * 16);
                uint32_t pk[8];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float4 bb = b4[j];
                    float t0 = v[4 * j] + bb.x, t1 = v[4 * j + 1] + bb.y, t2 = v[4 * j + 2] + bb.z, t3 = v[4 * j + 3] + bb.w;
                    if (p.lrelu) { t0 = fmaxf(t0, 0.2f * t0); t1 = fmaxf(t1, 0.2f * t1); t2 = fmaxf(t2, 0.2f * t2); t3 = fmaxf(t3, 0.2f * t3); }
                    __nv_bfloat162 h0 = __floats2bfloat162_rn(t0 * scale, t1 * scale), h1 = __floats2bfloat162_rn(t2 * scale, t3 * scale);
                    pk[2 * j] = *reinterpret_cast<uint32_t*>(&h0);
                    pk[2 * j + 1] = *reinterpret_cast<uint32_t*>(&h1);
                }
                if (valid) {
                    uint4* d4 = reinterpret_cast<uint4*>(dst + u * 16);
                    d4[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                    d4[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
                }
            }
            tc_fence_before();
            mbar_arrive(&tmem_empty[acc]);
            if (++acc == 2) { acc = 0; aph ^= 1u; }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == kMmaWarp) tmem_dealloc(tmem_base, p.tmem_cols);
}

// ------------------------------------------------------------------------------------------------
// weight packing: fp32 [Cout][Cin][3][3]  ->  bf16 [slice][tap][Cin/8][nt][8]
//   transpose_flip = 0 (fprop):  B[n = co][k = ci] of tap (ky,kx) = w[co][ci][ky][kx]
//   transpose_flip = 1 (dgrad):  the data gradient is a 3x3 convolution of dY with
//                                w'[ci][co][ky][kx] = w[co][ci][2-ky][2-kx]; n runs over ci, k over co.
//   `n_out`, `k_in` are the GEMM N and K channel counts of the packed operand.
// ------------------------------------------------------------------------------------------------
__global__ void k_pack_weights(const float* __restrict__ w, int Cout, int Cin, int transpose_flip, int Nt,
                               __nv_bfloat16* __restrict__ out) {
    const int n_out = transpose_flip ? Cin : Cout, k_in = transpose_flip ? Cout : Cin;
    const int nch = k_in >> 3;
    const int total = 9 * n_out * k_in;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        // destination order: slice, tap, chunk, n_local, e
        int r = i;
        const int e = r & 7; r >>= 3;
        // r indexes (slice, tap, chunk, n_local) with slice-dependent nt: decode by walking slices
        int slice = 0, base = 0;
        for (;;) {
            const int nt = min(Nt, n_out - slice * Nt);
            const int cnt = 9 * nch * nt;
            if (r < base + cnt) {
                const int q = r - base;
                const int n_local = q % nt, tc = q / nt;
                const int c = tc % nch, tap = tc / nch;
                const int n = slice * Nt + n_local, k = 8 * c + e;
                const int ky = tap / 3, kx = tap % 3;
                float v;
                if (!transpose_flip) v = w[(((size_t)n * Cin + k) * 3 + ky) * 3 + kx];
                else v = w[(((size_t)k * Cin + n) * 3 + (2 - ky)) * 3 + (2 - kx)];
                out[i] = __float2bfloat16_rn(v);
                break;
            }
            base += cnt; ++slice;
        }
    }
}

struct ConvPlan { int Nt, stages, tmem_cols, n_slices, occupancy, epi_warps; size_t smem; };

static ConvPlan plan_conv(int Cin, int Cout, bool need_full_n) {
    const size_t budget = 200 * 1024;
    const int nch = Cin / 8;
    ConvPlan pl{};
    for (int stages = 2; stages >= 1; --stages) {
        const size_t halo = (size_t)stages * nch * kHaloPitch * 16;
        for (int Nt = Cout; Nt >= 16; Nt -= 16) {
            const size_t wbytes = (size_t)9 * nch * Nt * 16;
            size_t tot = wbytes + halo + (size_t)(Nt + 8) * 4 + 1024 + 320;      // + PixelNorm partials + barriers
            if (tot <= budget) {
                const size_t stage_b = (size_t)nch * kHaloPitch * 16;
                int cols_ = 32; while (cols_ < 2 * Nt) cols_ <<= 1;
                // two CTAs per SM hide the latency of the (few, specialised) warps: take that shape when >= 4 halo slots
                // still fit in half of the shared memory and the TMEM columns of both fit
                // Several CTAs per SM hide the latency of the (few, specialised) warps.  Registers (72/thread) allow
                // 3 CTAs of 9 warps (4 epilogue warps) or 2 CTAs of 13 warps (8 epilogue warps); shared memory must hold
                // >= 4 halo slots per CTA and the TMEM columns of all resident CTAs must fit in 512.
                static const int force_occ = getenv("MG_CONV_OCC") ? atoi(getenv("MG_CONV_OCC")) : 0;
                const size_t base = tot - halo;
                pl.occupancy = 1; pl.epi_warps = 8;
                size_t cap = budget;
                if ((force_occ == 0 || force_occ == 3) && cols_ <= 128 && base + 4 * stage_b <= 73 * 1024) { pl.occupancy = 3; pl.epi_warps = 4; cap = 73 * 1024; }
                else if ((force_occ == 0 || force_occ >= 2) && cols_ <= 256 && base + 4 * stage_b <= 110 * 1024) { pl.occupancy = 2; cap = 110 * 1024; }
                if (pl.occupancy > 1) { tot = base + 4 * stage_b; stages = 4; }
                while (stages < kMaxStages && tot + stage_b <= cap) { ++stages; tot += stage_b; }
                pl.Nt = Nt; pl.stages = stages; pl.smem = tot;
                pl.n_slices = (Cout + Nt - 1) / Nt;
                int cols = 32; while (cols < 2 * Nt) cols <<= 1;
                pl.tmem_cols = cols;
                if (need_full_n && Nt != Cout) { pl.Nt = 0; }
                return pl;
            }
        }
    }
    pl.Nt = 0;
    return pl;
}

}  // namespace mg

using namespace mg;

extern "C" {

int mg_conv3x3_pack_weights(const float* w_f32, int Cin, int Cout, int dgrad, void* packed, size_t packed_bytes, mgStream stream) {
    if (!w_f32 || !packed) return MG_ERR_BAD_ARG;
    if (Cin < 16 || Cout < 16 || (Cin & 15) || (Cout & 15) || Cin > 256 || Cout > 256) return MG_ERR_UNSUPPORTED;
    if (packed_bytes < (size_t)9 * Cin * Cout * 2) return MG_ERR_WORKSPACE;
    ConvPlan pl = plan_conv(Cin, Cout, false);
    if (pl.Nt == 0) return MG_ERR_UNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    ProfScope ps("k_pack_weights", st);
    const int total = 9 * Cin * Cout;
    const int fwd_cout = dgrad ? Cin : Cout, fwd_cin = dgrad ? Cout : Cin;
    k_pack_weights<<<(total + 255) / 256, 256, 0, st>>>(w_f32, fwd_cout, fwd_cin, dgrad ? 1 : 0, pl.Nt, (__nv_bfloat16*)packed);
    return check_launch("k_pack_weights");
}

size_t mg_conv3x3_workspace_bytes(int Cin, int Cout) {
    return align_up((size_t)9 * Cin * Cout * 2, 256);
}

// flags: bit0 LeakyReLU(0.2), bit1 PixelNorm, bit2 input is nearest-upsampled x2 on the fly, bit3 dgrad weights
int mg_conv3x3_bf16(const void* x, const float* w_f32, const float* bias, void* y, float* inv_norm,
                    int B, int H, int W, int Cin, int Cout, int flags, void* ws, size_t ws_bytes, mgStream stream) {
    if (!x || !y || !ws) return MG_ERR_BAD_ARG;      // w_f32 == NULL: `ws` already holds the packed weights
    const bool dgrad = (flags & 8) != 0;
    // Cin/Cout are those of the GEMM actually run (for dgrad the caller passes Cin = channels of dY, Cout = channels of dX)
    if (B <= 0 || H <= 0 || W <= 0 || Cin < 16 || Cout < 16 || (Cin & 15) || (Cout & 15) || Cin > 256 || Cout > 256) return MG_ERR_UNSUPPORTED;
    const bool ups = (flags & 4) != 0;
    if (ups && ((H | W) & 1)) return MG_ERR_BAD_ARG;
    if (ws_bytes < mg_conv3x3_workspace_bytes(Cin, Cout)) return MG_ERR_WORKSPACE;
    const bool pn = (flags & 2) != 0;
    ConvPlan pl = plan_conv(Cin, Cout, pn);
    if (pl.Nt == 0) return MG_ERR_UNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    if (w_f32) {
        ProfScope ps("k_pack_weights", st);
        const int total = 9 * Cin * Cout;
        // w_f32 is [Cout_w][Cin_w][3][3] of the FORWARD convolution; for dgrad the GEMM's K (=Cin here) is the
        // forward Cout and the GEMM's N (=Cout here) the forward Cin
        const int fwd_cout = dgrad ? Cin : Cout, fwd_cin = dgrad ? Cout : Cin;
        k_pack_weights<<<(total + 255) / 256, 256, 0, st>>>(w_f32, fwd_cout, fwd_cin, dgrad ? 1 : 0, pl.Nt, (__nv_bfloat16*)ws);
    }
    ConvParams p{};
    p.x = (const __nv_bfloat16*)x; p.wpack = (const uint4*)ws; p.bias = bias; p.y = (__nv_bfloat16*)y; p.inv_norm = inv_norm;
    p.B = B; p.H = H; p.W = W; p.Hin = ups ? H / 2 : H; p.Win = ups ? W / 2 : W; p.Cin = Cin; p.Cout = Cout;
    p.upsample = ups; p.lrelu = flags & 1; p.pixelnorm = pn;
    p.tiles_x = (W + kTileW - 1) / kTileW; p.tiles_y = (H + kTileH - 1) / kTileH; p.n_tiles = B * p.tiles_x * p.tiles_y;
    p.Nt = pl.Nt; p.stages = pl.stages; p.tmem_cols = pl.tmem_cols; p.epi_warps = pl.epi_warps;
    p.idiv = make_item_div(Cin / 8);
    p.consumer_fence = getenv("MG_CONSUMER_FENCE") ? 1 : 0;
    p.tiles_per_img = p.tiles_x * p.tiles_y;
    p.div_img = make_fast_div(p.tiles_per_img);
    p.div_tx = make_fast_div(p.tiles_x);
    if (p.n_tiles >= (1 << 20)) return MG_ERR_UNSUPPORTED;
    static int sm_count = 0;
    if (!sm_count) { int dev; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev); }
    cudaFuncSetAttribute(k_conv3x3, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem);
    const int ctas = sm_count * pl.occupancy;
    const int per_slice = max(1, min(p.n_tiles, ctas / pl.n_slices > 0 ? ctas / pl.n_slices : 1));
    {
        ProfScope ps(dgrad ? "k_conv3x3_dgrad" : "k_conv3x3_fprop", st);
        k_conv3x3<<<dim3(per_slice, pl.n_slices), (pl.epi_warps + 5) * 32, pl.smem, st>>>(p);
    }
    return check_launch("k_conv3x3");
}

}  // extern "C"
