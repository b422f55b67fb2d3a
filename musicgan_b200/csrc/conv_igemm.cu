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
    float slope;                 // LeakyReLU as max(v, slope * v): 0.2, or 1.0 for the identity
    int tiles_x, tiles_y, n_tiles;
    int Nt, stages, tmem_cols, tiles_per_img;
    int n_acc, acc_stride;       // accumulators in flight in TMEM (2..8) and their column pitch
    int mb, blk_stride;          // M blocks (16 x 8 pixels each, stacked vertically) per pipeline step; TMEM columns per block
    int halo_pos, halo_pitch;    // (16 mb + 2) * 10 halo positions per step; pitch of one 8-channel plane of a slot
    ItemDiv idiv;
    FastDiv div_img, div_tx;
    int consumer_fence;
    int ablate;                  // debug (MG_CONV_ABLATE): 1 no halo copies, 2 no MMAs, 4 no epilogue work, 8 no global stores
    int epi_warps;               // 4 or 8 epilogue warps; producers are the next 4 warps, then the MMA warp
    int parts;                   // 1: bf16 weights; 2: weights as hi + lo bf16 pairs side by side along N: ONE MMA per tap and
                                 //    K step computes x * [w_hi | w_lo] into two column groups that the epilogue sums
};

// warp roles: [0, E) epilogue (E = 4, or 8 = two per TMEM lane quarter with half the columns each), [E, E+4) producers,
// warp E+4 MMA issue + TMEM alloc.  Block size (E + 5) * 32.
constexpr int kConvThreads = 13 * 32;                // upper bound (E = 8)
constexpr int kMaxAcc = 8;

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// kPN: PixelNorm in the epilogue.  kBA: bias + LeakyReLU in the epilogue (fprop); false for dgrad, whose epilogue is a
// plain fp32 -> bf16 store.  The bias is not added by the epilogue threads: it enters the accumulator through one extra
// K = 16 MMA per tile whose A operand is a constant tile (columns 0 and 1 = 1.0) and whose B operand holds the bias
// split in two bf16 terms (hi + lo: 16 mantissa bits) in K rows 0 and 1.
// kOcc4: the 9-warp shape compiled for 4 CTAs per SM (56 registers per thread, a shorter chunk table in the producers).
// kW2: hi + lo weights (two accumulator column groups summed by the epilogue); never combined with kOcc4 (registers).
template <bool kPN, bool kBA, bool kOcc4, bool kW2>
__global__ void __launch_bounds__(kOcc4 ? 288 : kConvThreads, kOcc4 ? 4 : 2)
k_conv3x3(const ConvParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid0 = threadIdx.x;
    const int nch = p.Cin >> 3;
    const int slice = blockIdx.y;
    const int n0 = slice * p.Nt;
    const int nt = min(p.Nt, p.Cout - n0);
    const int kEpiWarps = p.epi_warps, kProdWarp0 = kEpiWarps, kMmaWarp = kEpiWarps + 4;

    uint4* sW = reinterpret_cast<uint4*>(smem);                    // [9 taps][nch][parts][nt] + bias pseudo-tap [2][parts][nt]
    uint4* sOnes = sW + (9 * nch + 2) * p.parts * nt;              // [2][128]: the constant A tile of the bias MMA
    uint4* sA0 = sOnes + 256;
    float* sPart = reinterpret_cast<float*>(sA0 + (size_t)p.stages * nch * p.halo_pitch);      // [2][128] PixelNorm partial sums
    uint64_t* bars = reinterpret_cast<uint64_t*>(sPart + 256);
    uint64_t* full_a = bars;                       // [kMaxStages]
    uint64_t* empty_a = bars + kMaxStages;         // [kMaxStages]
    uint64_t* tmem_full = bars + 2 * kMaxStages;   // [kMaxAcc]
    uint64_t* tmem_empty = tmem_full + kMaxAcc;    // [kMaxAcc]
    uint64_t* w_full = tmem_full + 2 * kMaxAcc;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_full + 1);

    if (tid0 == 0) {
        for (int i = 0; i < kMaxStages; ++i) { mbar_init(&full_a[i], 128); mbar_init(&empty_a[i], 1); }
        for (int i = 0; i < kMaxAcc; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], kEpiWarps); }
        mbar_init(w_full, 256);           // per producer thread: its cp.async group + one plain (release) arrive
        mbar_fence_init();
    }
    if ((tid0 >> 5) == kMmaWarp) tmem_alloc(tmem_slot, p.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int acc_stride = p.acc_stride;
    // Thread index made opaque (bit 0 of a TMEM base address is always 0, but only at run time): ptxas otherwise
    // re-reads %tid.x (S2R, tens of cycles) inside every per-tile loop instead of keeping it in a register.
    const int tid = tid0 | (int)(tmem_base & 1u), warp = tid >> 5, lane = tid & 31;
    // Dependents may be scheduled only now that this CTA owns its TMEM columns: a dependent CTA that allocated first
    // would sit in pdl_wait() holding columns this CTA needs in order to finish -- a deadlock.
    pdl_trigger();
    pdl_wait();                  // everything above overlapped the previous kernel's tail; global memory from here on

    if (warp >= kProdWarp0 && warp < kMmaWarp) {
        // ================= producers =================
        const int pt = tid - kProdWarp0 * 32;
        {   // resident weights of this slice + bias
            const uint4* src = p.wpack + (size_t)9 * nch * n0 * p.parts;
            const int total = 9 * nch * nt * p.parts;
            // all chunks in flight at once (the deep layers carry up to 140 KB of weights per slice: a synchronous
            // load loop here used to dominate the run time of the small-spatial layers)
            const uint32_t sw_addr = smem_u32(sW);
            for (int i = pt; i < total; i += 128) cp_async16(sw_addr + (uint32_t)i * 16u, src + i, 16u);
            if (kBA) {
                uint4* sB = sW + 9 * nch * nt * p.parts;
                const int nn = p.parts * nt;                              // GEMM N: the bias sits in the first column group
                for (int i = pt; i < nn; i += 128) {
                    const float bv = (p.bias && i < nt) ? p.bias[n0 + i] : 0.0f;
                    const __nv_bfloat16 hi = __float2bfloat16_rn(bv), lo = __float2bfloat16_rn(bv - __bfloat162float(hi));
                    sB[i] = make_uint4((uint32_t)__bfloat16_as_ushort(hi) | ((uint32_t)__bfloat16_as_ushort(lo) << 16), 0u, 0u, 0u);
                    sB[nn + i] = make_uint4(0u, 0u, 0u, 0u);
                }
                sOnes[pt] = make_uint4(0x3F803F80u, 0u, 0u, 0u);          // bf16 (1.0, 1.0) in K columns 0 and 1
                sOnes[128 + pt] = make_uint4(0u, 0u, 0u, 0u);
                fence_proxy_async();                                      // st.shared operands -> tensor core reads
            }
            cp_async_arrive(w_full);
            mbar_arrive(w_full);
        }
        const int items = p.halo_pos * nch;
        // Per-thread item table, built once: for each 16-byte chunk this thread copies per tile, where it comes from
        // relative to the tile origin (tile independent, because tiles start at even rows / columns), where it goes in
        // the halo slot, and its halo coordinates (for the image-border test).  Up to kItemRegs items live in registers;
        // wider layers (Cin > 64) take the generic path for the remaining items.
        constexpr int kItemRegs = kOcc4 ? 6 : 12;
        int rel[kItemRegs];            // source offset in uint4 units from the tile origin
        uint32_t meta[kItemRegs];      // (hy << 25) | (hx << 21) | byte offset in the halo slot
        const int rows = kTileH * p.mb;
        const int sy_step = p.upsample ? rows / 2 : rows, sx_step = p.upsample ? kTileW / 2 : kTileW;
#pragma unroll
        for (int k = 0; k < kItemRegs; ++k) {
            const int i = pt + k * 128;
            rel[k] = 0; meta[k] = 0xFFFFFFFFu;
            if (i < items) {
                const int pos = (int)(((unsigned)i * p.idiv.magic) >> 20), c = i - pos * nch;
                const int hy = (pos * 6554) >> 16, hx = pos - hy * kHaloW;
                const int ry = p.upsample ? ((hy - 1) >> 1) : hy - 1, rx = p.upsample ? ((hx - 1) >> 1) : hx - 1;
                rel[k] = (ry * p.Win + rx) * nch + c;
                meta[k] = ((uint32_t)hy << 25) | ((uint32_t)hx << 21) | ((uint32_t)(c * p.halo_pitch + pos) * 16u);
            }
        }
        int slot = 0; uint32_t ph = 0;
        for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
            mbar_wait(&empty_a[slot], ph ^ 1u);
            const int b = fast_div(tile, p.div_img);
            const int tr = tile - b * p.tiles_per_img;
            const int tyi = fast_div(tr, p.div_tx), txi = tr - tyi * p.tiles_x;
            const int ty0 = tyi * rows - 1, tx0 = txi * kTileW - 1;
            const uint32_t dst = smem_u32(sA0 + (size_t)slot * nch * p.halo_pitch);
            const uint4* img = reinterpret_cast<const uint4*>(p.x + (size_t)b * p.Hin * p.Win * p.Cin);
            const uint4* org = img + ((size_t)(tyi * sy_step) * p.Win + txi * sx_step) * nch;
            // asynchronous 16-byte copies (zero fill outside the image); nothing is waited for here, so the loads of
            // up to `stages` tiles are in flight per CTA.  Tiles whose halo lies inside the image (the large majority at
            // the resolutions that dominate the run time) skip the per-chunk border logic.
            const bool interior = ty0 >= 0 && tx0 >= 0 && ty0 + rows + 2 <= p.H && tx0 + kTileW + 2 <= p.W;
            if (p.ablate & 1) {
            } else if (interior) {
#pragma unroll
                for (int k = 0; k < kItemRegs; ++k)
                    if (meta[k] != 0xFFFFFFFFu) cp_async16_full(dst + (meta[k] & 0x1FFFFFu), org + rel[k]);
            } else {
#pragma unroll
                for (int k = 0; k < kItemRegs; ++k) {
                    if (meta[k] != 0xFFFFFFFFu) {
                        const int iy = ty0 + (int)(meta[k] >> 25), ix = tx0 + (int)((meta[k] >> 21) & 0xF);
                        const bool ok = (unsigned)iy < (unsigned)p.H && (unsigned)ix < (unsigned)p.W;
                        cp_async16(dst + (meta[k] & 0x1FFFFFu), ok ? (const void*)(org + rel[k]) : (const void*)img, ok ? 16u : 0u);
                    }
                }
            }
            for (int i = pt + kItemRegs * 128; i < items; i += 128) {
                const int pos = (int)(((unsigned)i * p.idiv.magic) >> 20), c = i - pos * nch;
                const int hy = (pos * 6554) >> 16, hx = pos - hy * kHaloW;
                const int iy = ty0 + hy, ix = tx0 + hx;
                const bool ok = (unsigned)iy < (unsigned)p.H && (unsigned)ix < (unsigned)p.W;
                const int sy = p.upsample ? (iy >> 1) : iy, sx = p.upsample ? (ix >> 1) : ix;
                cp_async16(dst + (uint32_t)(c * p.halo_pitch + pos) * 16u,
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
        const int nn = p.parts * nt;                   // GEMM N of every MMA: [w | w_lo] side by side when parts == 2
        const uint32_t idesc = instr_desc_bf16(nn, false, false);
        // ablate bit 32 (timing experiment, wrong results): every tap reads 128-byte ALIGNED core matrices of slot 0
        const bool al = (p.ablate & 32) != 0;
        const uint64_t a_desc0 = al ? smem_desc((smem_u32(sA0) + 127u) & ~127u, 192u * 16u, 128u)
                                    : smem_desc(smem_u32(sA0), (uint32_t)p.halo_pitch * 16u, kHaloW * 16u);
        const uint64_t b_desc0 = smem_desc(smem_u32(sW), (uint32_t)nn * 16u, 128u);
        const uint64_t ones_desc = smem_desc(smem_u32(sOnes), 128u * 16u, 128u);
        const uint32_t slot_units = (uint32_t)(nch * p.halo_pitch), b_step = (uint32_t)(2 * nn);
        const uint32_t a_kstep = al ? 2u * 192u : 2u * (uint32_t)p.halo_pitch;
        const int ksteps = nch >> 1;
        mbar_wait(w_full, 0);
        int slot = 0, acc = 0; uint32_t ph = 0, aph = 0;
        for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
            mbar_wait(&tmem_empty[acc], aph ^ 1u);
            mbar_wait(&full_a[slot], ph);
            if (p.consumer_fence) fence_proxy_async();      // (debug switch) cp.async-written operands -> async proxy
            tc_fence_after();
            if (elect_one()) {       // single-thread region behind ONE elect.sync (see umma.cuh: `lane == 0` costs 4x per MMA)
                for (int blk = 0; blk < p.mb; ++blk) {
                    const uint32_t d = tmem_base + acc * acc_stride + blk * p.blk_stride;
                    const uint64_t da_blk = a_desc0 + (uint64_t)((al ? 0 : slot * slot_units) + blk * (kTileH * kHaloW));
                    uint64_t db = b_desc0;
                    uint32_t accum = 0;
                    // parts == 2 (w = hi + lo, 16 significand bits, so that LeakyReLU masks agree with the fp32
                    // reference's): the low halves sit beside the high halves along N, the SAME MMA computes both
                    // products.  (These small MMAs cost ~85 cycles each whatever N <= 64 -- the 4 KB activation tile is
                    // re-read from shared memory by every instruction -- and their count is what bounds the layer: a
                    // second sweep over the halo, as first written, cost 30-60 % on the 128 x 128 / 256 x 256 layers.)
                    if (!(p.ablate & 2))
#pragma unroll
                    for (int tap = 0; tap < 9; ++tap) {
                        uint64_t da = da_blk + (uint64_t)(al ? 0 : (tap / 3) * kHaloW + (tap % 3));
                        for (int kk = 0; kk < ksteps; ++kk) {
                            mma_bf16(d, da, db, idesc, accum);
                            accum = 1;
                            da += a_kstep;
                            db += b_step;
                        }
                    }
                    if (kBA) mma_bf16(d, ones_desc, db, idesc, 1u);      // + bias (db now points at the bias pseudo-tap)
                }
                // ONE commit per tile: the same mbarrier tells the producers that the halo slot is free again and the
                // epilogue warps that the accumulator is complete (a second tcgen05.commit per tile costs as much
                // tensor-pipe time as four of these small MMAs)
                mma_commit(&empty_a[slot]);
            }
            __syncwarp();
            if (++slot == p.stages) { slot = 0; ph ^= 1u; }
            if (++acc == p.n_acc) { acc = 0; aph ^= 1u; }
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
        int acc = 0, slot = 0; uint32_t aph = 0, ph = 0;
        for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
            const int b = fast_div(tile, p.div_img);
            const int tr = tile - b * p.tiles_per_img;
            const int tyi = fast_div(tr, p.div_tx), txi = tr - tyi * p.tiles_x;
            const int ox = txi * kTileW + rx;
            mbar_wait(&empty_a[slot], ph);            // k-th completion of this slot's barrier == MMAs of this step done
            tc_fence_after();
            // the accumulators go back to the MMA warp as soon as their last columns sit in registers, before the
            // arithmetic and the stores of those columns
            bool released = false;
            auto release = [&]() {
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tmem_empty[acc]);      // one arrival per epilogue warp
                released = true;
            };
            if (!(p.ablate & 4))
            for (int blk = 0; blk < p.mb; ++blk) {
                const int oy = (tyi * p.mb + blk) * kTileH + ry;
                const bool valid = oy < p.H && ox < p.W;
                const bool last_blk = blk + 1 == p.mb;
                const uint32_t taddr = tmem_base + acc * acc_stride + blk * p.blk_stride + ((uint32_t)(quarter * 32) << 16);
                // 16 accumulator columns of unit u; with hi + lo weights the sum of the two column groups
                auto load16 = [&](int u, float* v) {
                    tmem_ld16(taddr + u * 16, v);
                    if (kW2) {
                        float w16[16];
                        tmem_ld16(taddr + nt + u * 16, w16);
                        tmem_wait_ld();
#pragma unroll
                        for (int j = 0; j < 16; ++j) v[j] += w16[j];
                    }
                };
                float scale = 1.0f;
                if (kPN) {
                    float ss = 0.0f;
                    for (int u = u0; u < u1; ++u) {
                        float v[16];
                        load16(u, v);
                        tmem_wait_ld();
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const float t = kBA ? fmaxf(v[j], p.slope * v[j]) : v[j];
                            ss = fmaf(t, t, ss);
                        }
                    }
                    float tot = ss;
                    if (split) {
                        sPart[half * 128 + m] = ss;
                        named_bar_sync(1 + quarter, 64);              // the two warps that share this TMEM lane quarter
                        tot = sPart[m] + sPart[128 + m];
                        named_bar_sync(1 + quarter, 64);              // sPart may be overwritten by the next block
                    }
                    scale = 1.0f / sqrtf(tot * inv_c + 1e-8f);
                    if (valid && half == 0 && p.inv_norm) p.inv_norm[((size_t)b * p.H + oy) * p.W + ox] = scale;
                }
                __nv_bfloat16* dst = p.y + (((size_t)b * p.H + oy) * p.W + ox) * p.Cout + n0;
                // 16 channels: activation, PixelNorm scale, bf16 pack, two 16-byte stores
                auto finish16 = [&](const float* v, int u) {
                    uint32_t pk[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        float t0 = v[2 * j], t1 = v[2 * j + 1];
                        if (kBA) { t0 = fmaxf(t0, p.slope * t0); t1 = fmaxf(t1, p.slope * t1); }
                        if (kPN) { t0 *= scale; t1 *= scale; }
                        const __nv_bfloat162 h = __floats2bfloat162_rn(t0, t1);
                        pk[j] = *reinterpret_cast<const uint32_t*>(&h);
                    }
                    if (valid && (!(p.ablate & 8) || pk[0] == 0x12345678u)) {
                        // one 256-bit store = one full 32-byte sector per thread (two 16-byte stores reach L2 as two
                        // partial-sector writes)
                        asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                                     ::"l"(dst + u * 16), "r"(pk[0]), "r"(pk[1]), "r"(pk[2]), "r"(pk[3]),
                                       "r"(pk[4]), "r"(pk[5]), "r"(pk[6]), "r"(pk[7]) : "memory");
                    }
                };
                int u = u0;
                for (; u + 2 <= u1; u += 2) {
                    float v[32];
                    load16(u, v);
                    load16(u + 1, v + 16);
                    tmem_wait_ld();
                    if (last_blk && u + 2 == u1) release();
                    if (!(p.ablate & 16)) { finish16(v, u); finish16(v + 16, u + 1); }
                }
                if (u < u1) {
                    float v[16];
                    load16(u, v);
                    tmem_wait_ld();
                    if (last_blk) release();
                    if (!(p.ablate & 16)) finish16(v, u);
                }
            }
            if (!released) release();                 // a warp of the second half without columns of its own (Nt = 16)
            if (++acc == p.n_acc) { acc = 0; aph ^= 1u; }
            if (++slot == p.stages) { slot = 0; ph ^= 1u; }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == kMmaWarp) tmem_dealloc(tmem_base, p.tmem_cols);
}

// ------------------------------------------------------------------------------------------------
// weight packing: fp32 [Cout][Cin][3][3]  ->  bf16 [slice][tap][Cin/8][part][nt][8]   (part 0 = bf16(w), part 1 =
// bf16(w - part 0) when parts == 2: side by side along the GEMM N dimension)
//   transpose_flip = 0 (fprop):  B[n = co][k = ci] of tap (ky,kx) = w[co][ci][ky][kx]
//   transpose_flip = 1 (dgrad):  the data gradient is a 3x3 convolution of dY with
//                                w'[ci][co][ky][kx] = w[co][ci][2-ky][2-kx]; n runs over ci, k over co.
//   `n_out`, `k_in` are the GEMM N and K channel counts of the packed operand.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void pack_weights_range(const float* __restrict__ w, int Cout, int Cin, int transpose_flip, int Nt, int parts,
                                                   __nv_bfloat16* __restrict__ out, int first, int stride) {
    const int n_out = transpose_flip ? Cin : Cout, k_in = transpose_flip ? Cout : Cin;
    const int nch = k_in >> 3;
    const int total = 9 * n_out * k_in;
    for (int i = first; i < total; i += stride) {
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
                const __nv_bfloat16 hi = __float2bfloat16_rn(v);
                // slice base in the output is parts * base; within the slice [tap][chunk][part][n_local]
                const size_t o = ((size_t)parts * base + ((size_t)tc * parts) * nt + n_local) * 8 + e;
                out[o] = hi;
                if (parts == 2) out[o + (size_t)nt * 8] = __float2bfloat16_rn(v - __bfloat162float(hi));
                break;
            }
            base += cnt; ++slice;
        }
    }
}

__global__ void k_pack_weights(const float* __restrict__ w, int Cout, int Cin, int transpose_flip, int Nt, int parts,
                               __nv_bfloat16* __restrict__ out) {
    pdl_trigger();
    pdl_wait();
    pack_weights_range(w, Cout, Cin, transpose_flip, Nt, parts, out, blockIdx.x * blockDim.x + threadIdx.x, gridDim.x * blockDim.x);
}

// All packed copies of all weights of a network in ONE launch (a training step re-packs ~40 tensors after every
// optimiser step: as separate launches they cost more than the packing itself).  grid (blocks, n_jobs).
__global__ void __launch_bounds__(256)
k_pack_weights_multi(const mgPackJob* __restrict__ jobs) {
    pdl_trigger();
    pdl_wait();
    const mgPackJob j = jobs[blockIdx.y];
    const int first = blockIdx.x * blockDim.x + threadIdx.x, stride = gridDim.x * blockDim.x;
    if (j.kind == 0) pack_weights_range(j.w, j.cout_fwd, j.cin_fwd, j.flip, j.nt, j.parts, (__nv_bfloat16*)j.out, first, stride);
    else pack_weights_split_range(j.w, j.cout_fwd, j.cin_fwd, j.flip, j.parts, (__nv_bfloat16*)j.out, first, stride);
}

struct ConvPlan { int Nt, stages, tmem_cols, n_slices, occupancy, epi_warps, n_acc, acc_stride, mb, blk_stride, halo_pos, halo_pitch; size_t smem; };

static ConvPlan plan_conv(int Cin, int Cout, bool need_full_n, int H = 0, int parts = 1) {
    // hi + lo weights (parts == 2) double the resident weights: the PixelNorm layers (full N per CTA) then need the
    // whole 227 KB of a CTA and run with one or two halo slots
    const size_t budget = (parts == 2 && need_full_n ? 224 : 200) * 1024;
    const int nch = Cin / 8;
    ConvPlan pl{};
    // 1. N slice: the widest slice whose resident weights leave room for two single-block halo slots (one if need be).
    //    The packed weight layout depends on this choice only, never on the shape chosen in step 2.
    int Nt = 0;
    for (int stages = 2; stages >= 1 && !Nt; --stages)
        for (int n = Cout; n >= (need_full_n ? Cout : 16); n -= 16)
            if ((size_t)(9 * nch + 2) * n * 16 * parts + (size_t)stages * nch * kHaloPitch * 16 + 4096 + 1024 + 384 <= budget) { Nt = n; break; }
    if (!Nt || (need_full_n && Nt != Cout)) return pl;
    const size_t base = (size_t)(9 * nch + 2) * Nt * 16 * parts + 4096 + 1024 + 384;   // weights, bias pseudo-tap, constant A tile, PixelNorm partials, barriers
    // 2. shape of one pipeline step.  The warps of a CTA are few and specialised, so (a) several CTAs share an SM
    //    (registers: 72/thread -> 3 CTAs of 9 warps with 4 epilogue warps, or 2 CTAs of 13 warps with 8) and (b) on
    //    the large images a step covers `mb` vertically stacked 16 x 8 blocks, which divides the per-step costs
    //    (barrier round trips, tile index arithmetic, commit) by mb and shrinks the halo overhead.  Requirements: the
    //    per-thread chunk table of the producers stays in registers, >= 3 halo slots, >= 2 accumulators, and the
    //    TMEM columns of all resident CTAs fit in 512.
    const int force_occ = getenv("MG_CONV_OCC") ? atoi(getenv("MG_CONV_OCC")) : 0;
    const int force_mb = getenv("MG_CONV_MB") ? atoi(getenv("MG_CONV_MB")) : 0;
    const int force_acc = getenv("MG_CONV_ACC") ? atoi(getenv("MG_CONV_ACC")) : 0;
    static const int prefs[8][2] = {{2, 4}, {1, 4}, {2, 3}, {1, 3}, {4, 2}, {2, 2}, {1, 2}, {1, 1}};
    for (int pass = 0; pass < 2; ++pass) {            // pass 1 ignores the debug overrides if they leave no candidate
        for (const auto& c : prefs) {
            const int mb = c[0], occ = c[1];
            if (pass == 0 && ((force_occ && occ != force_occ) || (force_mb && mb != force_mb))) continue;
            if (mb > 1 && H < kTileH * mb) continue;
            if (occ == 4 && parts == 2) continue;       // the hi + lo epilogue does not fit the 56-register shape
            const int halo_pos = (kTileH * mb + 2) * kHaloW, pitch = halo_pos + 6;
            if (mb > 1 && halo_pos * nch > 12 * 128) continue;
            if (occ == 4 && halo_pos * nch > 6 * 128) continue;
            const size_t stage_b = (size_t)nch * pitch * 16;
            const size_t cap = occ == 4 ? 54 * 1024 : occ == 3 ? 73 * 1024 : occ == 2 ? 110 * 1024 : budget;
            const int cols_cap = occ >= 3 ? 128 : occ == 2 ? 256 : 512;
            const int blk_stride = (parts * Nt + 31) & ~31, acc_stride = mb * blk_stride;      // parts column groups per block
            const int min_stages = occ > 1 ? ((mb > 1 || occ == 4) ? 3 : 4) : 1;
            if (occ > 1 && 2 * acc_stride > cols_cap) continue;
            if (acc_stride > cols_cap || base + min_stages * stage_b > cap) continue;
            int stages = min_stages;
            size_t tot = base + stages * stage_b;
            while (stages < kMaxStages && tot + stage_b <= cap) { ++stages; tot += stage_b; }
            int n_acc = cols_cap / acc_stride;
            n_acc = n_acc > kMaxAcc ? kMaxAcc : n_acc;
            if (n_acc > stages) n_acc = stages;           // the shared slot barrier must not lap the epilogue (see kernel)
            if (force_acc >= 1 && force_acc <= n_acc) n_acc = force_acc;
            int cols = 32; while (cols < n_acc * acc_stride) cols <<= 1;
            pl.Nt = Nt; pl.stages = stages; pl.smem = tot; pl.n_slices = (Cout + Nt - 1) / Nt;
            pl.occupancy = occ; pl.epi_warps = occ >= 3 ? 4 : 8;
            pl.mb = mb; pl.blk_stride = blk_stride; pl.acc_stride = acc_stride; pl.n_acc = n_acc; pl.tmem_cols = cols;
            pl.halo_pos = halo_pos; pl.halo_pitch = pitch;
            return pl;
        }
    }
    return pl;
}

}  // namespace mg

using namespace mg;

extern "C" {

// mode: bit 0 = data-gradient orientation, bit 1 = hi + lo pairs (for mg_conv3x3_bf16 with flag 16), bit 2 = the layer
// runs with the PixelNorm epilogue (all output channels in one slice: the slice width is part of the packed layout)
int mg_conv3x3_pack_weights(const float* w_f32, int Cin, int Cout, int mode, void* packed, size_t packed_bytes, mgStream stream) {
    if (!w_f32 || !packed) return MG_ERR_BAD_ARG;
    if (Cin < 16 || Cout < 16 || (Cin & 15) || (Cout & 15) || Cin > 256 || Cout > 256) return MG_ERR_UNSUPPORTED;
    const int dgrad = mode & 1, parts = (mode & 2) ? 2 : 1;
    if (packed_bytes < (size_t)9 * Cin * Cout * 2 * parts) return MG_ERR_WORKSPACE;
    ConvPlan pl = plan_conv(Cin, Cout, (mode & 4) != 0, 0, parts);
    if (pl.Nt == 0) return MG_ERR_UNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    ProfScope ps("k_pack_weights", st);
    const int total = 9 * Cin * Cout;
    const int fwd_cout = dgrad ? Cin : Cout, fwd_cin = dgrad ? Cout : Cin;
    launch_pdl(k_pack_weights, dim3((total + 255) / 256), dim3(256), 0, st, w_f32, fwd_cout, fwd_cin, dgrad ? 1 : 0, pl.Nt, parts, (__nv_bfloat16*)packed);
    return check_launch("k_pack_weights");
}

// Fill `jobs[i]` (host memory) for one packed copy: kind 0 = mg_conv3x3_pack_weights(mode), kind 1 =
// mg_conv3x3_split_pack_weights(dgrad = mode & 1); Cin / Cout are those of the GEMM, as in the single-tensor calls.
int mg_pack_job_fill(mgPackJob* job, const float* w_f32, void* packed, int Cin, int Cout, int kind, int mode) {
    if (!job || !w_f32 || !packed) return MG_ERR_BAD_ARG;
    if (Cin < 16 || Cout < 16 || (Cin & 15) || (Cout & 15) || Cin > 256 || Cout > 256) return MG_ERR_UNSUPPORTED;
    const int dgrad = mode & 1;
    job->w = w_f32; job->out = packed; job->kind = kind ? 1 : 0; job->flip = dgrad;
    job->cout_fwd = dgrad ? Cin : Cout; job->cin_fwd = dgrad ? Cout : Cin;
    job->parts = kind ? ((mode & 2) ? 3 : 2) : ((mode & 2) ? 2 : 1);
    job->nt = 0;
    if (!kind) {
        ConvPlan pl = plan_conv(Cin, Cout, (mode & 4) != 0, 0, job->parts);
        if (pl.Nt == 0) return MG_ERR_UNSUPPORTED;
        job->nt = pl.Nt;
    }
    job->total = 9 * Cin * Cout;
    return MG_OK;
}

// `jobs_dev`: n_jobs mgPackJob records in DEVICE memory (filled on the host with mg_pack_job_fill and uploaded once: the
// pointers of parameters and packed buffers do not change between optimiser steps); max_total = largest job->total.
int mg_pack_weights_multi(const mgPackJob* jobs_dev, int n_jobs, int max_total, mgStream stream) {
    if (!jobs_dev || n_jobs <= 0 || max_total <= 0) return MG_ERR_BAD_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    ProfScope ps("k_pack_weights_multi", st);
    int blocks = (max_total + 256 * 8 - 1) / (256 * 8);
    if (blocks < 1) blocks = 1;
    launch_pdl(k_pack_weights_multi, dim3(blocks, n_jobs), dim3(256), 0, st, jobs_dev);
    return check_launch("k_pack_weights_multi");
}

size_t mg_conv3x3_workspace_bytes(int Cin, int Cout) {
    return align_up((size_t)9 * Cin * Cout * 4, 256);       // room for the hi + lo packing
}

// flags: bit0 LeakyReLU(0.2), bit1 PixelNorm, bit2 input is nearest-upsampled x2 on the fly, bit3 dgrad weights,
// bit4 weights as hi + lo bf16 pairs (two MMA sweeps)
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
    const int parts = (flags & 16) ? 2 : 1;
    ConvPlan pl = plan_conv(Cin, Cout, pn, H, parts);
    if (pl.Nt == 0) return MG_ERR_UNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    if (w_f32) {
        ProfScope ps("k_pack_weights", st);
        const int total = 9 * Cin * Cout;
        // w_f32 is [Cout_w][Cin_w][3][3] of the FORWARD convolution; for dgrad the GEMM's K (=Cin here) is the
        // forward Cout and the GEMM's N (=Cout here) the forward Cin
        const int fwd_cout = dgrad ? Cin : Cout, fwd_cin = dgrad ? Cout : Cin;
        launch_pdl(k_pack_weights, dim3((total + 255) / 256), dim3(256), 0, st, w_f32, fwd_cout, fwd_cin, dgrad ? 1 : 0, pl.Nt, parts, (__nv_bfloat16*)ws);
    }
    ConvParams p{};
    p.x = (const __nv_bfloat16*)x; p.wpack = (const uint4*)ws; p.bias = bias; p.y = (__nv_bfloat16*)y; p.inv_norm = inv_norm;
    p.B = B; p.H = H; p.W = W; p.Hin = ups ? H / 2 : H; p.Win = ups ? W / 2 : W; p.Cin = Cin; p.Cout = Cout;
    p.upsample = ups; p.lrelu = flags & 1; p.pixelnorm = pn;
    p.tiles_x = (W + kTileW - 1) / kTileW; p.tiles_y = (H + kTileH * pl.mb - 1) / (kTileH * pl.mb); p.n_tiles = B * p.tiles_x * p.tiles_y;
    p.Nt = pl.Nt; p.stages = pl.stages; p.tmem_cols = pl.tmem_cols; p.epi_warps = pl.epi_warps; p.parts = parts;
    p.n_acc = pl.n_acc; p.acc_stride = pl.acc_stride; p.mb = pl.mb; p.blk_stride = pl.blk_stride;
    p.halo_pos = pl.halo_pos; p.halo_pitch = pl.halo_pitch;
    p.idiv = make_item_div(Cin / 8);
    p.consumer_fence = getenv("MG_CONSUMER_FENCE") ? 1 : 0;
    p.ablate = getenv("MG_CONV_ABLATE") ? atoi(getenv("MG_CONV_ABLATE")) : 0;
    p.tiles_per_img = p.tiles_x * p.tiles_y;
    p.div_img = make_fast_div(p.tiles_per_img);
    p.div_tx = make_fast_div(p.tiles_x);
    if (p.n_tiles >= (1 << 20)) return MG_ERR_UNSUPPORTED;
    const int sm_count = current_sm_count();
    p.slope = (flags & 1) ? 0.2f : 1.0f;
    const bool ba = bias != nullptr || (flags & 1);
    const bool o4 = pl.occupancy == 4;
    auto kern = o4 ? (pn ? (ba ? k_conv3x3<true, true, true, false> : k_conv3x3<true, false, true, false>)
                         : (ba ? k_conv3x3<false, true, true, false> : k_conv3x3<false, false, true, false>))
                   : (pn ? (ba ? k_conv3x3<true, true, false, false> : k_conv3x3<true, false, false, false>)
                         : (ba ? k_conv3x3<false, true, false, false> : k_conv3x3<false, false, false, false>));
    if (parts == 2) {
        if (o4) return MG_ERR_UNSUPPORTED;
        kern = pn ? (ba ? k_conv3x3<true, true, false, true> : k_conv3x3<true, false, false, true>)
                  : (ba ? k_conv3x3<false, true, false, true> : k_conv3x3<false, false, false, true>);
    }
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem);
    int ctas = sm_count * pl.occupancy;
    if (const char* e = getenv("MG_CONV_MAX_CTAS")) { const int m = atoi(e); if (m > 0 && m < ctas) ctas = m; }   // tests: many steps per CTA
    const int per_slice = max(1, min(p.n_tiles, ctas / pl.n_slices > 0 ? ctas / pl.n_slices : 1));
    {
        ProfScope ps(dgrad ? "k_conv3x3_dgrad" : "k_conv3x3_fprop", st);
        launch_pdl(kern, dim3(per_slice, pl.n_slices), dim3((pl.epi_warps + 5) * 32), pl.smem, st, p);
    }
    return check_launch("k_conv3x3");
}

}  // extern "C"
