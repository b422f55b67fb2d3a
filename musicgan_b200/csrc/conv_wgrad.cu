// Weight gradient of the 3x3 / s1 / p1 convolutions as a tcgen05 GEMM whose reduction runs over PIXELS.
//
//   dw[co][ci][ky][kx] = sum_{b,y,x} dy[b][y][x][co] * xin[b][y+ky-1][x+kx-1][ci]
//
// (autograd's convolution_backward weight term for the nn.Conv2d of reference networks/generator.py:16-37 and
// networks/discriminator.py:15-32).  Per 16 x 8 pixel tile both operands are staged once in the canonical
// no-swizzle layout [channel chunk][pixel][8 ch] and read as MN-major UMMA operands (umma.cuh):
//   A = dy tile   : M = co (8-channel chunks, SBO = chunk pitch), K = 128 tile pixels (dense, LBO = 128 B)
//   B = xin halo  : N = ci,                                       K = the same pixels shifted by the tap:
//                   16 pixels of one MMA = two tile rows -> start (2j+ky)*10+kx, LBO = one halo row (160 B)
// so per tile 8 MMAs (M128 x N=Cin x K16) per tap accumulate D_tap[co][ci] in TMEM; the accumulators of all
// tiles of the CTA stay in TMEM (9 taps x Cin columns, split in tap groups when that exceeds 512 columns) and
// are flushed ONCE at the end with red.global.add.f32.
#include "common.cuh"
#include "umma.cuh"
#include "conv_common.cuh"
#include <cstdlib>

namespace mg {
using namespace umma;

constexpr int kDyPitch = 129;          // pixels per channel chunk of the staged dy tile (odd: spreads the transposer's banks)
constexpr int kWgradMmaWarps = 3;      // taps are dealt round-robin to three MMA-issuing warps (72 MMAs per tile)
constexpr int kXposeWarps = 8;         // dy tile: smem -> registers -> TMEM (A operand), two warps per TMEM lane quarter
constexpr int kWgradThreads = (kXposeWarps + 4 + kWgradMmaWarps) * 32;     // 480

struct WgradParams {
    const __nv_bfloat16* dy;    // [B][H][W][Cout]
    const __nv_bfloat16* x;     // [B][Hin][Win][Cin]
    float* dw;                  // [Cout][Cin][3][3], accumulated atomically
    int B, H, W, Hin, Win, Cin, Cout, upsample;
    int tiles_x, tiles_y, n_tiles;
    int taps_per_group, tmem_cols, stages, dy_chunks;
    ItemDiv idiv_x;
    unsigned bar_offset;
    int tiles_per_img;
    FastDiv div_img, div_tx;
};

// Why the A operand lives in TMEM: with both operands in shared memory every one of the 72 MMAs of a tile re-reads the
// 128-row (mostly padding) dy operand, 4 KB each -> ~290 KB of shared-memory reads per tile, which bounds the kernel
// at ~2500 cycles/tile.  Here the dy tile is transposed ONCE per tile into TMEM (lane = output channel, column = pixel
// pair) and all nine taps read it from there; shared memory only serves the small shifted x windows.
// Output channel co sits on TMEM lane (co & 3) * 32 + (co >> 2) so that the four lane quarters (and therefore the
// transposer warps) share the work evenly even when Cout is small.
__global__ void __launch_bounds__(kWgradThreads, 1)
k_conv3x3_wgrad(const WgradParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nch_x = p.Cin >> 3;
    const int tap0 = blockIdx.y * p.taps_per_group;
    const int ntaps = min(p.taps_per_group, 9 - tap0);
    const int co0 = blockIdx.z * 128;
    const int co_n = min(128, p.Cout - co0);
    const int nch_dy = co_n >> 3;

    const size_t dy_bytes = (size_t)p.dy_chunks * kDyPitch * 16;
    const size_t x_bytes = (size_t)nch_x * kHaloPitch * 16;
    const size_t stage_bytes = dy_bytes + x_bytes;
    unsigned char* stage0 = smem;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.bar_offset);
    uint64_t* full = bars;                    // [kMaxStages]  producers -> transposers + MMA
    uint64_t* empty = bars + kMaxStages;      // [kMaxStages]  MMA commit -> producers
    uint64_t* a_full = bars + 2 * kMaxStages; // [2]           transposers -> MMA
    uint64_t* a_empty = a_full + 2;           // [2]           MMA commit -> transposers
    uint64_t* done = a_full + 4;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a_full + 5);

    constexpr int kProd0 = kXposeWarps, kMma0 = kXposeWarps + 4;
    if (tid == 0) {
        for (int i = 0; i < kMaxStages; ++i) { mbar_init(&full[i], 128); mbar_init(&empty[i], kWgradMmaWarps); }
        for (int i = 0; i < 2; ++i) { mbar_init(&a_full[i], kXposeWarps * 32); mbar_init(&a_empty[i], kWgradMmaWarps); }
        mbar_init(done, kWgradMmaWarps);
        mbar_fence_init();
    }
    if (warp == kMma0) tmem_alloc(tmem_slot, p.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_a = tmem_base + (uint32_t)(p.tmem_cols - 128);      // two A buffers of 64 columns

    if (warp >= kProd0 && warp < kMma0) {
        // ================= producers =================
        const int pt = tid - kProd0 * 32;
        const unsigned dy_magic = ((1u << 20) + nch_dy - 1) / nch_dy;
        int slot = 0; uint32_t ph = 0;
        for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
            mbar_wait(&empty[slot], ph ^ 1u);
            const int b = fast_div(tile, p.div_img);
            const int tr = tile - b * p.tiles_per_img;
            const int tyi = fast_div(tr, p.div_tx);
            const int oy0 = tyi * kTileH, ox0 = (tr - tyi * p.tiles_x) * kTileW;
            const uint32_t s_dy = smem_u32(stage0 + slot * stage_bytes);
            const uint32_t s_x = s_dy + (uint32_t)dy_bytes;
            // dy tile: 128 pixels x nch_dy chunks (zero outside the image: those pixels must not contribute)
            const __nv_bfloat16* dyb = p.dy + (size_t)b * p.H * p.W * p.Cout + co0;
            const int items_dy = 128 * nch_dy;
            for (int i = pt; i < items_dy; i += 128) {
                const int pix = (int)(((unsigned)i * dy_magic) >> 20), c = i - pix * nch_dy;
                const int oy = oy0 + (pix >> 3), ox = ox0 + (pix & 7);
                const bool ok = oy < p.H && ox < p.W;
                const void* src = ok ? (const void*)(reinterpret_cast<const uint4*>(dyb + ((size_t)oy * p.W + ox) * p.Cout) + c) : (const void*)p.dy;
                cp_async16(s_dy + (uint32_t)(c * kDyPitch + pix) * 16u, src, ok ? 16u : 0u);
            }
            // x halo (same staging as the forward kernel)
            const __nv_bfloat16* xb = p.x + (size_t)b * p.Hin * p.Win * p.Cin;
            const int items_x = kHaloPos * nch_x;
            for (int i = pt; i < items_x; i += 128) {
                const int pos = (int)(((unsigned)i * p.idiv_x.magic) >> 20), c = i - pos * nch_x;
                const int hy = (pos * 6554) >> 16, hx = pos - hy * kHaloW;
                const int iy = oy0 - 1 + hy, ix = ox0 - 1 + hx;
                const bool ok = (unsigned)iy < (unsigned)p.H && (unsigned)ix < (unsigned)p.W;
                const int sy = p.upsample ? (iy >> 1) : iy, sx = p.upsample ? (ix >> 1) : ix;
                const void* src = ok ? (const void*)(reinterpret_cast<const uint4*>(xb + ((size_t)sy * p.Win + sx) * p.Cin) + c) : (const void*)p.x;
                cp_async16(s_x + (uint32_t)(c * kHaloPitch + pos) * 16u, src, ok ? 16u : 0u);
            }
            cp_async_arrive(&full[slot]);
            if (++slot == p.stages) { slot = 0; ph ^= 1u; }
        }
    } else if (warp >= kMma0) {
        // ================= MMA issue: warp kMma0 + w issues taps w, w + 3, w + 6 (A from TMEM, B from smem) =================
        const int mw = warp - kMma0;
        const uint32_t idesc = instr_desc_bf16(p.Cin, false, true);
        const uint64_t b_desc0 = smem_desc(smem_u32(stage0) + (uint32_t)dy_bytes, kHaloW * 16u, kHaloPitch * 16u);
        const uint32_t stage_units = (uint32_t)(stage_bytes >> 4);
        int slot = 0, buf = 0; uint32_t ph = 0, aph = 0, accum = 0;
        for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
            mbar_wait(&full[slot], ph);
            mbar_wait(&a_full[buf], aph);
            tc_fence_after();
            if (lane == 0) {
                const uint64_t db0 = b_desc0 + (uint64_t)(slot * stage_units);
                const uint32_t a0 = tmem_a + buf * 64;
                for (int tl = mw; tl < ntaps; tl += kWgradMmaWarps) {
                    const int tap = tap0 + tl;
                    const int ky = tap / 3, kx = tap - ky * 3;
                    const uint32_t d = tmem_base + tl * p.Cin;
                    uint64_t db = db0 + (uint64_t)(ky * kHaloW + kx);
                    mma_bf16_ts(d, a0, db, idesc, accum);
#pragma unroll
                    for (int j = 1; j < 8; ++j) {
                        db += 2 * kHaloW;
                        mma_bf16_ts(d, a0 + j * 8, db, idesc, 1u);
                    }
                }
                mma_commit(&empty[slot]);
                mma_commit(&a_empty[buf]);
            }
            accum = 1;
            __syncwarp();
            if (++slot == p.stages) { slot = 0; ph ^= 1u; }
            if (++buf == 2) { buf = 0; aph ^= 1u; }
        }
        if (lane == 0) mma_commit(done);
        __syncwarp();
    } else {
        // ================= transposers: dy tile smem -> TMEM A operand; then (warps 0-3) the final flush =================
        const int quarter = warp & 3, khalf = warp >> 2;
        const int co = 4 * lane + quarter;                  // local output channel owned by this TMEM lane
        const bool row_ok = co < co_n;
        const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
        int slot = 0, buf = 0; uint32_t ph = 0, aph = 0;
        for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
            mbar_wait(&a_empty[buf], aph ^ 1u);
            mbar_wait(&full[slot], ph);
            tc_fence_after();
            const unsigned char* s_dy = stage0 + slot * stage_bytes;
            const unsigned char* row = s_dy + ((size_t)(co >> 3) * kDyPitch) * 16 + (co & 7) * 2;
            const uint32_t ta = tmem_a + buf * 64 + khalf * 32 + lane_addr;
#pragma unroll
            for (int g = 0; g < 4; ++g) {                    // 4 groups of 8 columns = 16 pixels each
                uint32_t r[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int pix = khalf * 64 + g * 16 + 2 * j;
                    uint32_t lo = 0, hi = 0;
                    if (row_ok) {
                        lo = *reinterpret_cast<const uint16_t*>(row + (size_t)pix * 16);
                        hi = *reinterpret_cast<const uint16_t*>(row + (size_t)(pix + 1) * 16);
                    }
                    r[j] = lo | (hi << 16);
                }
                tmem_st8(ta + g * 8, r);
            }
            tmem_wait_st();
            tc_fence_before();
            mbar_arrive(&a_full[buf]);
            if (++slot == p.stages) { slot = 0; ph ^= 1u; }
            if (++buf == 2) { buf = 0; aph ^= 1u; }
        }
        {   // final flush: the two warps of a lane quarter split the taps; plain stores when no other CTA contributes
            mbar_wait(done, 0);
            tc_fence_after();
            const uint32_t taddr = tmem_base + lane_addr;
            const bool exclusive = gridDim.x == 1;
            for (int tl = khalf; tl < ntaps; tl += 2) {
                const int tap = tap0 + tl;
                for (int c0 = 0; c0 < p.Cin; c0 += 16) {
                    float v[16];
                    tmem_ld16(taddr + tl * p.Cin + c0, v);
                    tmem_wait_ld();
                    if (row_ok) {
                        float* dst = p.dw + ((size_t)(co0 + co) * p.Cin + c0) * 9 + tap;
                        if (exclusive) {
#pragma unroll
                            for (int j = 0; j < 16; ++j) dst[j * 9] = v[j];
                        } else {
#pragma unroll
                            for (int j = 0; j < 16; ++j) atomicAdd(dst + j * 9, v[j]);
                        }
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == kMma0) tmem_dealloc(tmem_base, p.tmem_cols);
}

}  // namespace mg

using namespace mg;

extern "C" int mg_conv3x3_wgrad_bf16(const void* dy, const void* x, float* dw, float* dbias_unused,
                                     int B, int H, int W, int Cin, int Cout, int upsample_in, mgStream stream) {
    (void)dbias_unused;
    if (!dy || !x || !dw) return MG_ERR_BAD_ARG;
    if (B <= 0 || H <= 0 || W <= 0 || Cin < 16 || Cout < 16 || (Cin & 15) || (Cout & 15) || Cin > 256 || Cout > 256) return MG_ERR_UNSUPPORTED;
    if (upsample_in && ((H | W) & 1)) return MG_ERR_BAD_ARG;
    WgradParams p{};
    p.dy = (const __nv_bfloat16*)dy; p.x = (const __nv_bfloat16*)x; p.dw = dw;
    p.B = B; p.H = H; p.W = W; p.Hin = upsample_in ? H / 2 : H; p.Win = upsample_in ? W / 2 : W; p.Cin = Cin; p.Cout = Cout;
    p.upsample = upsample_in ? 1 : 0;
    p.tiles_x = (W + kTileW - 1) / kTileW; p.tiles_y = (H + kTileH - 1) / kTileH; p.n_tiles = B * p.tiles_x * p.tiles_y;
    // Tap groups: a CTA keeps taps_per_group x Cin accumulator columns plus two 64-column A buffers in its 512 TMEM
    // columns; more taps than fit are split over blockIdx.y (balanced, e.g. 9 -> 5 + 4), each group re-reading the tiles.
    int tpp = (512 - 128) / Cin; tpp = tpp > 9 ? 9 : (tpp < 1 ? 1 : tpp);
    int groups = (9 + tpp - 1) / tpp;
    static int sm_count_ = 0;
    if (!sm_count_) { int dev; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sm_count_, cudaDevAttrMultiProcessorCount, dev); }
    {   // small-spatial layers: too few tiles to fill the GPU -> spread the taps over more CTAs instead (fewer columns to
        // flush per CTA, and with one CTA per group no atomics at all)
        const int mb = (Cout + 127) / 128;
        while (groups < 9 && p.n_tiles * groups * mb < sm_count_) groups = groups < 3 ? 3 : 9;
    }
    p.taps_per_group = (9 + groups - 1) / groups;
    groups = (9 + p.taps_per_group - 1) / p.taps_per_group;
    int cols = 32; while (cols < p.taps_per_group * Cin + 128) cols <<= 1;
    p.tmem_cols = cols;
    const int occ = 1;
    const int mblocks = (Cout + 127) / 128;
    p.dy_chunks = (mblocks > 1 ? 128 : Cout) / 8;           // chunks of the widest M block
    const size_t stage_bytes = (size_t)p.dy_chunks * kDyPitch * 16 + (size_t)(Cin / 8) * kHaloPitch * 16;
    int stages = kMaxStages;
    size_t smem = 0;
    for (; stages >= 1; --stages) {
        const size_t need = (size_t)stages * stage_bytes;
        smem = align_up(need, 16) + 256;
        if (smem <= 200 * 1024) { p.bar_offset = (unsigned)align_up(need, 16); break; }
    }
    if (stages < 1) return MG_ERR_UNSUPPORTED;
    p.stages = stages;
    p.idiv_x = make_item_div(Cin / 8);
    p.tiles_per_img = p.tiles_x * p.tiles_y;
    p.div_img = make_fast_div(p.tiles_per_img);
    p.div_tx = make_fast_div(p.tiles_x);
    if (p.n_tiles >= (1 << 20)) return MG_ERR_UNSUPPORTED;
    static int sm_count = 0;
    if (!sm_count) { int dev; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev); }
    cudaFuncSetAttribute(k_conv3x3_wgrad, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int gx = sm_count * occ / (groups * mblocks);
    gx = gx < 1 ? 1 : gx;
    gx = gx > p.n_tiles ? p.n_tiles : gx;
    cudaStream_t st = (cudaStream_t)stream;
    {
        ProfScope ps("k_conv3x3_wgrad", st);
        k_conv3x3_wgrad<<<dim3(gx, groups, mblocks), kWgradThreads, smem, st>>>(p);
    }
    return check_launch("k_conv3x3_wgrad");
}
