// Weight gradient of the 3x3 / s1 / p1 convolutions as a tcgen05 GEMM whose reduction runs over PIXELS.
//
//   dw[co][ci][ky][kx] = sum_{b,y,x} dy[b][y][x][co] * xin[b][y+ky-1][x+kx-1][ci]
//
// (autograd's convolution_backward weight term for the nn.Conv2d of reference networks/generator.py:16-37 and
// networks/discriminator.py:15-32).  GEMM view per 16 x 8 pixel tile (K = its 128 pixels):
//
//   D[(tap, ci)][co]  +=  A[(tap, ci)][k] * B[k][co]        A[(tap,ci)][k] = xin[pixel k shifted by tap][ci]
//                                                           B[k][co]       = dy[pixel k][co]
//
// * the nine taps are STACKED IN M: rows r = tap * Cin + ci, cut in blocks of 128 rows, so the 128 rows of an MMA are
//   full even when the channel counts are 16..48 (with M = output channels, as first written, 1/8..3/8 of the rows were
//   used and 72 tiny MMAs per tile were issue bound);
// * A lives in TMEM: transposer warps read the staged x halo ([channel chunk][halo position][8 ch], the same staging as
//   the forward kernel) and write, per block, lane = row, column = pixel pair with tcgen05.st; the MMA takes A from
//   TMEM (TS mode) and B = the dy tile MN-major from shared memory (no-swizzle canonical layout, umma.cuh);
// * per tile and block 8 MMAs (M128 x N=Cout x K16); the accumulators of ALL tiles of a CTA stay in TMEM and are flushed
//   once at the end into a per-CTA slice of the workspace; k_wgrad_reduce sums the slices into dw (deterministic, no
//   atomics: 148 CTAs adding 9*Cin*Cout values each onto the same addresses cost 20-30 us per launch).
#include "common.cuh"
#include "umma.cuh"
#include "conv_common.cuh"
#include <cstdlib>

namespace mg {
using namespace umma;

constexpr int kDyPitch = 129;          // pixels per channel chunk of the staged dy tile (B operand)
constexpr int kWgradMmaWarps = 3;      // row blocks are dealt round-robin to three MMA-issuing warps (one warp: D0.conv1 76 -> 85 us)
constexpr int kXposeWarps = 8;         // x halo: smem -> registers -> TMEM (A operand), two warps per TMEM lane quarter
constexpr int kWgradThreads = (kXposeWarps + 4 + kWgradMmaWarps) * 32;     // 480
constexpr int kMaxBlocks = 4;          // row blocks per CTA (TMEM: nb * (Cout + a_bufs * 64) columns)
constexpr int kMaxABufs = 4;           // A-operand buffers in TMEM (transposers run that many tiles ahead of the MMAs)

struct WgradParams {
    const __nv_bfloat16* dy;    // [B][H][W][Cout]
    const __nv_bfloat16* x;     // [B][Hin][Win][Cin]
    float* part;                // per-CTA partial sums [gridDim.x][groups][nb][128 rows][Cout] (workspace)
    int B, H, W, Hin, Win, Cin, Cout, upsample;
    int tiles_x, tiles_y, n_tiles;
    int blocks_per_cta, n_blocks, tmem_cols, stages;
    ItemDiv idiv_x, idiv_dy;
    FastDiv div_cin;
    unsigned bar_offset;
    int tiles_per_img;
    FastDiv div_img, div_tx;
    int ablate;                 // debug (MG_WGRAD_ABLATE): 1 no copies, 2 no MMAs, 4 no operand staging (ldmatrix / tcgen05.st)
    int a_bufs;                 // A-operand buffers in TMEM (2 .. kMaxABufs, as many as the 512 columns allow)
    int bias_row;               // 9 * Cin: an A row of ones whose products are the bias gradient sum_pixels dy[.][co]; -1 = none
};

__global__ void __launch_bounds__(kWgradThreads, 1)
k_conv3x3_wgrad(const WgradParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid0 = threadIdx.x;
    const int nch_x = p.Cin >> 3, nch_dy = p.Cout >> 3;
    const int blk0 = blockIdx.y * p.blocks_per_cta;
    const int nb = min(p.blocks_per_cta, p.n_blocks - blk0);
    const int rows_total = 9 * p.Cin;

    const size_t dy_bytes = (size_t)nch_dy * kDyPitch * 16;
    const size_t x_bytes = (size_t)nch_x * kHaloPitch * 16;
    const size_t stage_bytes = dy_bytes + x_bytes;
    unsigned char* stage0 = smem;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.bar_offset);
    uint64_t* full = bars;                    // [kMaxStages]  producers -> transposers + MMA
    uint64_t* empty = bars + kMaxStages;      // [kMaxStages]  MMA commit -> producers
    uint64_t* a_full = bars + 2 * kMaxStages; // [kMaxABufs]   transposers -> MMA
    uint64_t* a_empty = a_full + kMaxABufs;   // [kMaxABufs]   MMA commit -> transposers
    uint64_t* done = a_full + 2 * kMaxABufs;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a_full + 2 * kMaxABufs + 1);

    constexpr int kProd0 = kXposeWarps, kMma0 = kXposeWarps + 4;
    if (tid0 == 0) {
        for (int i = 0; i < kMaxStages; ++i) { mbar_init(&full[i], (p.ablate & 8) ? 4 : 128); mbar_init(&empty[i], kWgradMmaWarps + kXposeWarps); }
        for (int i = 0; i < kMaxABufs; ++i) { mbar_init(&a_full[i], kXposeWarps); mbar_init(&a_empty[i], kWgradMmaWarps); }
        mbar_init(done, kWgradMmaWarps);
        mbar_fence_init();
    }
    if ((tid0 >> 5) == kMma0) tmem_alloc(tmem_slot, p.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // Thread index made opaque (bit 0 of a TMEM base address is always 0, but only at run time): ptxas otherwise
    // re-reads %tid.x (S2R, tens of cycles) inside every per-tile loop instead of keeping it in a register.
    const int tid = tid0 | (int)(tmem_base & 1u), warp = tid >> 5, lane = tid & 31;
    const int a_cols = p.blocks_per_cta * 64;                                  // columns of one A buffer
    const uint32_t tmem_a = tmem_base + (uint32_t)(p.tmem_cols - p.a_bufs * a_cols);  // the A buffers at the top
    pdl_trigger();               // only after the TMEM allocation (see k_conv3x3: a dependent must not allocate first)
    pdl_wait();                  // everything above overlapped the previous kernel's tail; global memory from here on

    if (warp >= kProd0 && warp < kMma0) {
        // ================= producers: dy tile + x halo of the tile, 16-byte cp.async =================
        const int pt = tid - kProd0 * 32;
        const int items_dy = 128 * nch_dy, items_x = kHaloPos * nch_x;
        // Per-thread chunk tables, built once (tile independent because tiles start at even rows / columns): source
        // offset from the tile origin in 16-byte units, and (y << 25) | (x << 21) | byte offset in the slot.  Tiles whose
        // halo lies inside the image copy without any per-chunk border logic.  Layers wider than 64 channels take the
        // generic loop for the chunks beyond the tables.
        constexpr int kDyRegs = 8, kXRegs = 12;
        int rel_dy[kDyRegs], rel_x[kXRegs];
        uint32_t meta_dy[kDyRegs], meta_x[kXRegs];
#pragma unroll
        for (int k = 0; k < kDyRegs; ++k) {
            const int i = pt + k * 128;
            rel_dy[k] = 0; meta_dy[k] = 0xFFFFFFFFu;
            if (i < items_dy) {
                const int pix = (int)(((unsigned)i * p.idiv_dy.magic) >> 20), c = i - pix * nch_dy;
                rel_dy[k] = ((pix >> 3) * p.W + (pix & 7)) * nch_dy + c;
                meta_dy[k] = ((uint32_t)(pix >> 3) << 25) | ((uint32_t)(pix & 7) << 21) | ((uint32_t)(c * kDyPitch + pix) * 16u);
            }
        }
#pragma unroll
        for (int k = 0; k < kXRegs; ++k) {
            const int i = pt + k * 128;
            rel_x[k] = 0; meta_x[k] = 0xFFFFFFFFu;
            if (i < items_x) {
                const int pos = (int)(((unsigned)i * p.idiv_x.magic) >> 20), c = i - pos * nch_x;
                const int hy = (pos * 6554) >> 16, hx = pos - hy * kHaloW;
                const int ry = p.upsample ? ((hy - 1) >> 1) : hy - 1, rx = p.upsample ? ((hx - 1) >> 1) : hx - 1;
                rel_x[k] = (ry * p.Win + rx) * nch_x + c;
                meta_x[k] = ((uint32_t)hy << 25) | ((uint32_t)hx << 21) | ((uint32_t)(c * kHaloPitch + pos) * 16u);
            }
        }
        int slot = 0; uint32_t ph = 0;
        for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
            mbar_wait(&empty[slot], ph ^ 1u);
            const int b = fast_div(tile, p.div_img);
            const int tr = tile - b * p.tiles_per_img;
            const int tyi = fast_div(tr, p.div_tx);
            const int oy0 = tyi * kTileH, ox0 = (tr - tyi * p.tiles_x) * kTileW;
            const uint32_t s_dy = smem_u32(stage0 + slot * stage_bytes);
            const uint32_t s_x = s_dy + (uint32_t)dy_bytes;
            const __nv_bfloat16* dyb = p.dy + (size_t)b * p.H * p.W * p.Cout;
            const __nv_bfloat16* xb = p.x + (size_t)b * p.Hin * p.Win * p.Cin;
            const uint4* org_dy = reinterpret_cast<const uint4*>(dyb + ((size_t)oy0 * p.W + ox0) * p.Cout);
            const uint4* org_x = reinterpret_cast<const uint4*>(xb + ((size_t)(p.upsample ? oy0 >> 1 : oy0) * p.Win + (p.upsample ? ox0 >> 1 : ox0)) * p.Cin);
            const bool interior = oy0 >= 1 && ox0 >= 1 && oy0 + kTileH + 1 <= p.H && ox0 + kTileW + 1 <= p.W;
            if (p.ablate & 1) {
            } else if (interior) {
#pragma unroll
                for (int k = 0; k < kDyRegs; ++k)
                    if (meta_dy[k] != 0xFFFFFFFFu) cp_async16_full(s_dy + (meta_dy[k] & 0x1FFFFFu), org_dy + rel_dy[k]);
#pragma unroll
                for (int k = 0; k < kXRegs; ++k)
                    if (meta_x[k] != 0xFFFFFFFFu) cp_async16_full(s_x + (meta_x[k] & 0x1FFFFFu), org_x + rel_x[k]);
            } else {
                // dy tile: zero outside the image (those pixels must not contribute); x halo: zero padding
#pragma unroll
                for (int k = 0; k < kDyRegs; ++k) {
                    if (meta_dy[k] != 0xFFFFFFFFu) {
                        const bool ok = oy0 + (int)(meta_dy[k] >> 25) < p.H && ox0 + (int)((meta_dy[k] >> 21) & 0xF) < p.W;
                        cp_async16(s_dy + (meta_dy[k] & 0x1FFFFFu), ok ? (const void*)(org_dy + rel_dy[k]) : (const void*)p.dy, ok ? 16u : 0u);
                    }
                }
#pragma unroll
                for (int k = 0; k < kXRegs; ++k) {
                    if (meta_x[k] != 0xFFFFFFFFu) {
                        const int iy = oy0 - 1 + (int)(meta_x[k] >> 25), ix = ox0 - 1 + (int)((meta_x[k] >> 21) & 0xF);
                        const bool ok = (unsigned)iy < (unsigned)p.H && (unsigned)ix < (unsigned)p.W;
                        cp_async16(s_x + (meta_x[k] & 0x1FFFFFu), ok ? (const void*)(org_x + rel_x[k]) : (const void*)p.x, ok ? 16u : 0u);
                    }
                }
            }
            for (int i = pt + kDyRegs * 128; i < items_dy; i += 128) {
                const int pix = (int)(((unsigned)i * p.idiv_dy.magic) >> 20), c = i - pix * nch_dy;
                const int oy = oy0 + (pix >> 3), ox = ox0 + (pix & 7);
                const bool ok = oy < p.H && ox < p.W;
                const void* src = ok ? (const void*)(reinterpret_cast<const uint4*>(dyb + ((size_t)oy * p.W + ox) * p.Cout) + c) : (const void*)p.dy;
                cp_async16(s_dy + (uint32_t)(c * kDyPitch + pix) * 16u, src, ok ? 16u : 0u);
            }
            for (int i = pt + kXRegs * 128; i < items_x; i += 128) {
                const int pos = (int)(((unsigned)i * p.idiv_x.magic) >> 20), c = i - pos * nch_x;
                const int hy = (pos * 6554) >> 16, hx = pos - hy * kHaloW;
                const int iy = oy0 - 1 + hy, ix = ox0 - 1 + hx;
                const bool ok = (unsigned)iy < (unsigned)p.H && (unsigned)ix < (unsigned)p.W;
                const int sy = p.upsample ? (iy >> 1) : iy, sx = p.upsample ? (ix >> 1) : ix;
                const void* src = ok ? (const void*)(reinterpret_cast<const uint4*>(xb + ((size_t)sy * p.Win + sx) * p.Cin) + c) : (const void*)p.x;
                cp_async16(s_x + (uint32_t)(c * kHaloPitch + pos) * 16u, src, ok ? 16u : 0u);
            }
            if (p.ablate & 8) { if (lane == 0) mbar_arrive(&full[slot]); }      // timing experiment: 4 arrivals instead of 128
            else cp_async_arrive(&full[slot]);
            if (++slot == p.stages) { slot = 0; ph ^= 1u; }
        }
    } else if (warp >= kMma0) {
        // ================= MMA issue: warp kMma0 + w issues row blocks w, w + 3 (A from TMEM, B = dy from smem) =================
        const int mw = warp - kMma0;
        const uint32_t idesc = instr_desc_bf16(p.Cout, false, true);
        const uint64_t b_desc0 = smem_desc(smem_u32(stage0), 128u, kDyPitch * 16u);    // MN-major: 8-pixel groups 128 B apart
        const uint32_t stage_units = (uint32_t)(stage_bytes >> 4);
        int slot = 0, buf = 0; uint32_t ph = 0, aph = 0, accum = 0;
        for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
            mbar_wait(&full[slot], ph);
            mbar_wait(&a_full[buf], aph);
            tc_fence_after();
            if (elect_one()) {       // single-thread region behind ONE elect.sync (umma.cuh)
                const uint64_t db0 = b_desc0 + (uint64_t)(slot * stage_units);
                for (int j = mw; j < ((p.ablate & 2) ? 0 : nb); j += kWgradMmaWarps) {
                    const uint32_t d = tmem_base + j * p.Cout;
                    const uint32_t a0 = tmem_a + buf * a_cols + j * 64;
                    uint64_t db = db0;
                    mma_bf16_ts(d, a0, db, idesc, accum);
#pragma unroll
                    for (int s = 1; s < 8; ++s) {
                        db += 16;                              // next 16 pixels of the dy tile
                        mma_bf16_ts(d, a0 + s * 8, db, idesc, 1u);
                    }
                }
                mma_commit(&empty[slot]);
                mma_commit(&a_empty[buf]);
            }
            accum = 1;
            __syncwarp();
            if (++slot == p.stages) { slot = 0; ph ^= 1u; }
            if (++buf == p.a_bufs) { buf = 0; aph ^= 1u; }
        }
        if (elect_one()) mma_commit(done);      // the lane that issued the MMAs
        __syncwarp();
    } else {
        // ================= transposers: shifted x windows, smem -> TMEM A operand; then the final flush =================
        const int quarter = warp & 3, khalf = warp >> 2;
        const int L = quarter * 32 + lane;                   // TMEM lane == row within a block
        const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
        // A 32-row quarter of a block is two groups of 16 rows = one tap x two 8-channel chunks (Cin is a multiple of
        // 16).  Per group and pair of tile rows (y, y+1) ONE ldmatrix.x4.trans fetches four 8 x 8 matrices
        //     (chunk c0, row y) (chunk c0+1, row y) (chunk c0, row y+1) (chunk c0+1, row y+1)
        // -- 8 pixels x 8 channels each, 128 contiguous bytes of the staged halo -- and hands thread i, per matrix, the
        // pixel pair 2(i%4), 2(i%4)+1 of channel i/4 packed in one register: exactly the fragment that
        // tcgen05.st.16x128b.x2 scatters to TMEM lanes i/4 and i/4+8, columns i%4 and 4+i%4 (probed:
        // mg_debug_tmem_store), i.e. lane = row (tap, ci), column = pixel pair.  Two instructions move 16 rows x 16
        // pixels; the former per-thread path needed 16 two-byte loads and 8 byte-permutes for ONE row.
        int goff[kMaxBlocks][2];       // per block and 16-row group: byte offset of (chunk c0, tap) in a staged x halo, or -1 (padding rows)
        const int lane_part = (((lane >> 3) & 1) * kHaloPitch + (lane >> 4) * kHaloW + (lane & 7)) * 16;
#pragma unroll
        for (int j = 0; j < kMaxBlocks; ++j) {
#pragma unroll
            for (int g = 0; g < 2; ++g) {
                const int r0 = (blk0 + j) * 128 + quarter * 32 + g * 16;
                goff[j][g] = -1;
                if (j < nb && r0 < rows_total) {
                    const int tap = fast_div(r0, p.div_cin), ci0 = r0 - tap * p.Cin;
                    const int ky = tap / 3, kx = tap - ky * 3;
                    goff[j][g] = ((ci0 >> 3) * kHaloPitch + ky * kHaloW + kx) * 16 + lane_part;
                }
            }
        }
        // Bias gradient: row 9 * Cin of the A operand is a row of ones, so D[9 * Cin][co] = sum over the pixels of dy[.][co]
        // (dy is zero outside the image).  9 * Cin is a multiple of 16: the row opens a 16-row group that no tile writes
        // (goff < 0), so it is written ONCE here into both A buffers; its neighbours in the lane quarter get zeros (the
        // valid ones are rewritten for every tile, the others only feed accumulator rows that nobody reads).
        if (p.bias_row >= 0) {
            const int jb = (p.bias_row >> 7) - blk0;
            if (jb >= 0 && jb < nb && ((p.bias_row & 127) >> 5) == quarter) {
                uint32_t r[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) r[k] = L == (p.bias_row & 127) ? 0x3F803F80u : 0u;
                for (int b2 = 0; b2 < p.a_bufs; ++b2)
                    for (int c = 0; c < 4; ++c)
                        tmem_st8(tmem_a + b2 * a_cols + jb * 64 + khalf * 32 + c * 8 + lane_addr, r);
                tmem_wait_st();
            }
        }
        int slot = 0, buf = 0; uint32_t ph = 0, aph = 0;
        for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
            mbar_wait(&a_empty[buf], aph ^ 1u);
            mbar_wait(&full[slot], ph);
            tc_fence_after();
            const uint32_t s_x = smem_u32(stage0 + slot * stage_bytes + dy_bytes) + (uint32_t)(khalf * 8 * kHaloW * 16);
#pragma unroll
            for (int j = 0; j < kMaxBlocks; ++j) {
#pragma unroll
                for (int g = 0; g < 2; ++g) {
                    if (goff[j][g] >= 0 && !(p.ablate & 4)) {
                        const uint32_t ta = tmem_a + buf * a_cols + j * 64 + khalf * 32 + lane_addr + ((uint32_t)(g * 16) << 16);
                        const uint32_t src = s_x + (uint32_t)goff[j][g];
#pragma unroll
                        for (int t = 0; t < 4; ++t) {            // tile rows 8*khalf + 2t, +1  ->  columns 8t .. 8t+7
                            uint32_t r0, r1, r2, r3;
                            ldmatrix_x4_trans(src + (uint32_t)(t * 2 * kHaloW * 16), r0, r1, r2, r3);
                            tmem_st_16x128b_x2(ta + t * 8, r0, r1, r2, r3);
                        }
                    }
                }
            }
            tmem_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(&a_full[buf]);
                mbar_arrive(&empty[slot]);   // this warp no longer reads the smem slot (the MMA warps commit theirs)
            }
            if (++slot == p.stages) { slot = 0; ph ^= 1u; }
            if (++buf == p.a_bufs) { buf = 0; aph ^= 1u; }
        }
        {   // final flush: every CTA writes its accumulators ONCE, with plain 16-byte stores, to its own slice of the
            // workspace (k_wgrad_reduce sums the slices: no atomics, deterministic order, no pre-zeroed output)
            mbar_wait(done, 0);
            tc_fence_after();
            float* mine = p.part + ((size_t)(blockIdx.x * gridDim.y + blockIdx.y) * p.blocks_per_cta) * 128 * p.Cout;
#pragma unroll
            for (int j = 0; j < kMaxBlocks; ++j) {
                if (j < nb && (j & 1) == khalf) {
                    const uint32_t taddr = tmem_base + lane_addr + j * p.Cout;
                    float4* dst = reinterpret_cast<float4*>(mine + ((size_t)j * 128 + L) * p.Cout);
                    for (int c0 = 0; c0 < p.Cout; c0 += 16) {
                        float v[16];
                        tmem_ld16(taddr + c0, v);
                        tmem_wait_ld();
#pragma unroll
                        for (int q = 0; q < 4; ++q) dst[(c0 >> 2) + q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == kMma0) tmem_dealloc(tmem_base, p.tmem_cols);
}

// dw[co][ci][tap] = sum over the gx CTAs of a group of part[cta][group][j][L][co], rows r = tap*Cin + ci = (g*nb + j)*128 + L;
// with `db` the row r = 9 * Cin holds the bias gradient.  accumulate: bit 0 add to dw, bit 1 add to db (else overwrite).
// block = 8 warps x 32 consecutive outputs (consecutive co -> 128-byte rows); warp w sums CTAs w, w+8, ...; smem combine
__global__ void __launch_bounds__(256)
k_wgrad_reduce(const float* __restrict__ part, float* __restrict__ dw, float* __restrict__ db, int Cin, int Cout, int gx, int groups, int nb,
               FastDiv div_cin, int accumulate) {
    __shared__ float red[8][33];
    pdl_trigger();
    pdl_wait();
    const int total = 9 * Cin * Cout, total_all = total + (db ? Cout : 0);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int t = blockIdx.x * 32 + lane;
    float s0 = 0.f, s1 = 0.f;
    int r = 0, co = 0;
    if (t < total_all) {
        if (t < total) { r = t / Cout; co = t - r * Cout; } else { r = 9 * Cin; co = t - total; }
        const int blk = r >> 7, L = r & 127;
        const int g = blk / nb, j = blk - g * nb;
        const float* src = part + (((size_t)g * nb + j) * 128 + L) * Cout + co;
        const size_t cta_stride = (size_t)groups * nb * 128 * Cout;
        int x = w;
        for (; x + 8 < gx; x += 16) { s0 += src[(size_t)x * cta_stride]; s1 += src[(size_t)(x + 8) * cta_stride]; }
        if (x < gx) s0 += src[(size_t)x * cta_stride];
    }
    red[w][lane] = s0 + s1;
    __syncthreads();
    if (w == 0 && t < total_all) {
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) s += red[k][lane];
        if (t < total) {
            const int tap = fast_div(r, div_cin), ci = r - tap * Cin;
            float* dst = dw + ((size_t)co * Cin + ci) * 9 + tap;
            *dst = (accumulate & 1) ? *dst + s : s;      // a further contribution to the same parameter (ops.WgradLane)
        } else {
            db[co] = (accumulate & 2) ? db[co] + s : s;
        }
    }
}

}  // namespace mg

using namespace mg;

static void wgrad_grid(int n_tiles, int Cin, int Cout, int with_bias, int* nb_out, int* groups_out, int* gx_out) {
    const int sm_count = current_sm_count();
    const int n_blocks = (9 * Cin + (with_bias ? 1 : 0) + 127) / 128;
    int nb = 512 / (Cout + 128); nb = nb > kMaxBlocks ? kMaxBlocks : (nb < 1 ? 1 : nb);
    nb = nb > n_blocks ? n_blocks : nb;
    while (nb > 1 && n_tiles * ((n_blocks + nb - 1) / nb) < sm_count) --nb;
    int groups = (n_blocks + nb - 1) / nb;
    nb = (n_blocks + groups - 1) / groups;                 // balance
    groups = (n_blocks + nb - 1) / nb;
    int gx = sm_count / groups;
    gx = gx < 1 ? 1 : gx;
    gx = gx > n_tiles ? n_tiles : gx;
    *nb_out = nb; *groups_out = groups; *gx_out = gx;
}

extern "C" size_t mg_conv3x3_wgrad_workspace_bytes(int B, int H, int W, int Cin, int Cout) {
    if (B <= 0 || H <= 0 || W <= 0 || Cin < 16 || Cout < 16) return 0;
    const int n_tiles = B * ((W + kTileW - 1) / kTileW) * ((H + kTileH - 1) / kTileH);
    size_t need = 0;
    for (int with_bias = 0; with_bias < 2; ++with_bias) {      // one size for both entry points
        int nb, groups, gx;
        wgrad_grid(n_tiles, Cin, Cout, with_bias, &nb, &groups, &gx);
        const size_t n = (size_t)gx * groups * nb * 128 * Cout * sizeof(float);
        need = n > need ? n : need;
    }
    return align_up(need, 256);
}

extern "C" int mg_conv3x3_wgrad_bias_bf16(const void* dy, const void* x, float* dw, float* db, void* ws, size_t ws_bytes,
                                          int B, int H, int W, int Cin, int Cout, int flags, mgStream stream) {
    const int upsample_in = flags & 1, accumulate = (flags >> 1) & 3;
    if (!dy || !x || !dw || !ws) return MG_ERR_BAD_ARG;
    if (ws_bytes < mg_conv3x3_wgrad_workspace_bytes(B, H, W, Cin, Cout)) return MG_ERR_WORKSPACE;
    if (((uintptr_t)ws & 15) != 0) return MG_ERR_BAD_ARG;
    if (B <= 0 || H <= 0 || W <= 0 || Cin < 16 || Cout < 16 || (Cin & 15) || (Cout & 15) || Cin > 256 || Cout > 256) return MG_ERR_UNSUPPORTED;
    if (upsample_in && ((H | W) & 1)) return MG_ERR_BAD_ARG;
    WgradParams p{};
    p.dy = (const __nv_bfloat16*)dy; p.x = (const __nv_bfloat16*)x; p.part = (float*)ws;
    p.B = B; p.H = H; p.W = W; p.Hin = upsample_in ? H / 2 : H; p.Win = upsample_in ? W / 2 : W; p.Cin = Cin; p.Cout = Cout;
    p.upsample = upsample_in ? 1 : 0;
    p.tiles_x = (W + kTileW - 1) / kTileW; p.tiles_y = (H + kTileH - 1) / kTileH; p.n_tiles = B * p.tiles_x * p.tiles_y;
    // Row blocks: rows r = tap * Cin + ci in blocks of 128.  A CTA keeps nb * Cout accumulator columns plus two A buffers
    // of nb * 64 columns in its 512 TMEM columns; more blocks than fit are split over blockIdx.y (each group re-reads
    // the tiles).  When there are few tiles the blocks are spread over more CTAs anyway (small-spatial layers).
    p.bias_row = db ? 9 * Cin : -1;
    p.n_blocks = (9 * Cin + (db ? 1 : 0) + 127) / 128;
    int nb, groups, gx;
    wgrad_grid(p.n_tiles, Cin, Cout, db ? 1 : 0, &nb, &groups, &gx);
    p.blocks_per_cta = nb;
    int cols = 32; while (cols < nb * (Cout + 128)) cols <<= 1;
    p.tmem_cols = cols;
    {   // spare columns become further A buffers: the operand staging then runs more tiles ahead of the MMAs
        const int forced = getenv("MG_WGRAD_ABUFS") ? atoi(getenv("MG_WGRAD_ABUFS")) : 0;
        int a_bufs = (cols - nb * Cout) / (nb * 64);
        a_bufs = a_bufs > kMaxABufs ? kMaxABufs : a_bufs;
        if (forced >= 2 && forced <= a_bufs) a_bufs = forced;
        p.a_bufs = a_bufs;
    }
    const int occ = 1;
    const size_t stage_bytes = (size_t)(Cout / 8) * kDyPitch * 16 + (size_t)(Cin / 8) * kHaloPitch * 16;
    int stages = kMaxStages;
    size_t smem = 0;
    for (; stages >= 1; --stages) {
        const size_t need = (size_t)stages * stage_bytes;
        smem = align_up(need, 16) + 320;       // barriers: 2 * kMaxStages + 2 * kMaxABufs + 1, TMEM slot
        if (smem <= 200 * 1024) { p.bar_offset = (unsigned)align_up(need, 16); break; }
    }
    if (stages < 1) return MG_ERR_UNSUPPORTED;
    p.stages = stages;
    p.ablate = getenv("MG_WGRAD_ABLATE") ? atoi(getenv("MG_WGRAD_ABLATE")) : 0;
    p.idiv_x = make_item_div(Cin / 8);
    p.idiv_dy = make_item_div(Cout / 8);
    p.div_cin = make_fast_div(Cin);
    p.tiles_per_img = p.tiles_x * p.tiles_y;
    p.div_img = make_fast_div(p.tiles_per_img);
    p.div_tx = make_fast_div(p.tiles_x);
    if (p.n_tiles >= (1 << 20)) return MG_ERR_UNSUPPORTED;
    cudaFuncSetAttribute(k_conv3x3_wgrad, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaStream_t st = (cudaStream_t)stream;
    {
        ProfScope ps("k_conv3x3_wgrad", st);
        launch_pdl(k_conv3x3_wgrad, dim3(gx, groups, 1), dim3(kWgradThreads), smem, st, p);
    }
    {
        ProfScope ps("k_wgrad_reduce", st);
        const int total = 9 * Cin * Cout + (db ? Cout : 0);
        launch_pdl(k_wgrad_reduce, dim3((total + 31) / 32), dim3(256), 0, st, (const float*)ws, dw, db, Cin, Cout, gx, groups, nb, p.div_cin, accumulate);
    }
    return check_launch("k_conv3x3_wgrad");
}

extern "C" int mg_conv3x3_wgrad_bf16(const void* dy, const void* x, float* dw, void* ws, size_t ws_bytes,
                                     int B, int H, int W, int Cin, int Cout, int flags, mgStream stream) {
    return mg_conv3x3_wgrad_bias_bf16(dy, x, dw, nullptr, ws, ws_bytes, B, H, W, Cin, Cout, flags & 3, stream);
}
