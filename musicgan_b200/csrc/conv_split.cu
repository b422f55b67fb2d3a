// 3x3 / stride 1 / pad 1 convolution on fp32 NHWC activations with SPLIT-bf16 operands (sm_100a tcgen05):
// the precise path of the low-resolution layers (output height <= 32) of the ProGAN blocks
// (reference music_gan/networks/generator.py:16-22,31-37, networks/discriminator.py:15-21,26-32).
//
// Why: the critic is piecewise linear (LeakyReLU); a pre-activation whose sign flips under operand rounding changes
// the gradient discontinuously, and on the small-spatial layers (few pixels, nothing averages out) bf16 operands
// (2^-9) move the WGAN-GP gradients by 2..10 % (scripts/precision_study.py).  Here every operand is the sum of two
// bf16 numbers, x = hi + lo (16 significand bits), and the product is accumulated in fp32 as
//       hi_x * hi_w  +  lo_x * hi_w  +  hi_x * lo_w
// which puts the pre-activation error near 2^-16: the masks agree with the fp32 reference's.  The critic's forward
// convolutions go one step further (flag 32): w = hi + mid + lo carries the fp32 weight exactly and two more products
// (lo_x * mid_w, hi_x * lo_w) are added -- with 1e5 pre-activations per step a few sit within 1e-5 of zero, and one
// flipped mask in a 2 x 2-pixel layer moves the gradient-penalty gradient by percents (golden case stage2_b3).
// The weight parts of a K chunk sit side by side in shared memory, so the products of one activation half are ONE
// tcgen05.mma with GEMM N = parts * slice:  hi_x * [hi_w | mid_w | lo_w]  and  lo_x * [hi_w | mid_w]  -- two MMAs per
// (tap, 16 channels) instead of three or five, into `parts` column groups of one TMEM accumulator that the epilogue
// sums.  (Ablation, round 2: an M128 x N x K16 MMA fed from shared memory costs ~85 cycles whatever N <= 64 -- the
// 4 KB activation tile is re-read by every instruction -- and those MMAs were 2/3 of the run time of the small layers;
// separate accumulators per term did not help: the cost is issue throughput, not dependency latency.)
//
// Structure (small layers: simplicity over the last cycle):
//   * one CTA = one 16 x 8 pixel tile of one image (GEMM M = 128) x one slice of output channels;
//   * the reduction is streamed in groups of 16 input channels through a ring of shared-memory stages: per stage
//     4 producer warps read the 18 x 10 halo of those channels in fp32 and split it on the fly into the hi / lo bf16
//     planes of the canonical no-swizzle UMMA layout (umma.cuh) -- zero padding, image borders and the nearest x2
//     upsampling of the input are resolved in the read -- while the 4 (until then idle) epilogue warps copy the packed
//     weight parts of the group with cp.async.  (ncu, round 2: with ONE warp per scheduler doing both, the kernel was
//     bound by the dependent-instruction latency of those warps -- issue slots 22 % busy, tensor pipe 12-22 %.)
//   * one thread issues 9 taps x 3 MMAs (M128 x N x K16) per stage -- the taps are nine start addresses into the
//     halo -- and commits the stage back to the producers;
//   * 4 epilogue warps: tcgen05.ld -> + bias -> LeakyReLU(0.2) -> PixelNorm (layers.py:11-17) -> fp32 NHWC store.
// Data gradient = the same kernel on flipped / transposed packed weights, without bias / activation.
#include "common.cuh"
#include "umma.cuh"
#include "conv_common.cuh"
#include <cstdlib>

namespace mg {
using namespace umma;

constexpr int kSplitStagesMax = 6;
constexpr int kSplitThreads = 9 * 32;            // 4 epilogue warps, 4 producer warps, 1 MMA warp
constexpr int kSplitAPlane = 2 * kHaloPitch * 16; // bytes of one (hi or lo) plane of a stage: [2 chunks][186 pos][16 B]

struct SplitParams {
    const float* x;              // [B][Hin][Win][Cin] fp32
    const uint4* wpack;          // [Cin/16][9 taps][2 chunks][wparts][Cout][8 bf16]
    int wparts;                  // 2: w = hi + lo; 3: w = hi + mid + lo
    int ablate;                  // debug (MG_SPLIT_ABLATE): 1 no halo loads / conversion, 2 no weight copies, 4 no MMAs, 8 no stores
    const float* bias;           // [Cout] or null
    float* y;                    // [B][H][W][Cout] fp32
    __nv_bfloat16* y16;          // optional bf16 copy of y (operand of the bf16 weight-gradient kernel), or null
    float* inv_norm;             // [B][H][W] (pixelnorm, optional)
    int B, H, W, Hin, Win, Cin, Cout;
    int upsample, lrelu, pixelnorm;
    int tiles_x, tiles_y, Nt, stages, tmem_cols;
    unsigned stage_bytes;
};

__device__ __forceinline__ void split8(const float4 a, const float4 b, uint4& hi, uint4& lo) {
    const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    uint32_t h[4], l[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const __nv_bfloat16 h0 = __float2bfloat16_rn(v[2 * j]), h1 = __float2bfloat16_rn(v[2 * j + 1]);
        const __nv_bfloat16 l0 = __float2bfloat16_rn(v[2 * j] - __bfloat162float(h0));
        const __nv_bfloat16 l1 = __float2bfloat16_rn(v[2 * j + 1] - __bfloat162float(h1));
        h[j] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
        l[j] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
    }
    hi = make_uint4(h[0], h[1], h[2], h[3]);
    lo = make_uint4(l[0], l[1], l[2], l[3]);
}

template <bool kPN, bool kBA>
__global__ void __launch_bounds__(kSplitThreads, 2)
k_conv3x3_split(const SplitParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n0 = blockIdx.y * p.Nt;
    const int nt = min(p.Nt, p.Cout - n0);
    const int n_cg = p.Cin >> 4;

    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * p.stage_bytes);
    uint64_t* full = bars;                          // [stages] producers -> MMA
    uint64_t* empty = bars + kSplitStagesMax;       // [stages] MMA commit -> producers
    uint64_t* acc_full = bars + 2 * kSplitStagesMax;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);

    if (tid == 0) {
        for (int i = 0; i < kSplitStagesMax; ++i) { mbar_init(&full[i], 256); mbar_init(&empty[i], 1); }
        mbar_init(acc_full, 1);
        mbar_fence_init();
    }
    if (warp == 8) tmem_alloc(tmem_slot, p.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_trigger();               // after the TMEM allocation (see k_conv3x3)
    pdl_wait();

    const int tile = blockIdx.x;
    const int tiles_per_img = p.tiles_x * p.tiles_y;
    const int b = tile / tiles_per_img, tr = tile - b * tiles_per_img;
    const int tyi = tr / p.tiles_x, txi = tr - tyi * p.tiles_x;

    if (warp >= 4 && warp < 8) {
        // ================= producers =================
        const int pt = tid - 128;
        const int ty0 = tyi * kTileH - 1, tx0 = txi * kTileW - 1;
        const float* img = p.x + (size_t)b * p.Hin * p.Win * p.Cin;
        // this thread's (up to three) halo items: 8 channels of one halo position each
        int src_off[3]; uint32_t dst_off[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const int it = pt + k * 128;
            src_off[k] = -2; dst_off[k] = 0;
            if (it < 2 * kHaloPos) {
                const int pos = it >> 1, c = it & 1;
                const int hy = pos / kHaloW, hx = pos - hy * kHaloW;
                const int iy = ty0 + hy, ix = tx0 + hx;
                dst_off[k] = (uint32_t)(c * kHaloPitch + pos) * 16u;
                if ((unsigned)iy < (unsigned)p.H && (unsigned)ix < (unsigned)p.W) {
                    const int sy = p.upsample ? (iy >> 1) : iy, sx = p.upsample ? (ix >> 1) : ix;
                    src_off[k] = (sy * p.Win + sx) * p.Cin + c * 8;
                } else {
                    src_off[k] = -1;                 // padding: zeros
                }
            }
        }
        // The halo values are read with plain loads (they are converted before they reach shared memory), so a thread
        // that loaded, converted and stored one channel group after the other would pay one memory round trip per group
        // (measured: 2-3 us per group, 15-20 us for a 160-channel layer whatever its size).  The loads of the next
        // kPrefetch groups are therefore kept in flight in registers while the current group is converted.
        constexpr int kPrefetch = 2;
        float4 va[kPrefetch][3], vb[kPrefetch][3];
        auto load_group = [&](int cg, float4 (&a)[3], float4 (&b4)[3]) {
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                a[k] = b4[k] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (src_off[k] >= 0 && !(p.ablate & 1)) {
                    const float4* sp = reinterpret_cast<const float4*>(img + src_off[k] + cg * 16);
                    a[k] = __ldg(sp); b4[k] = __ldg(sp + 1);
                }
            }
        };
#pragma unroll
        for (int d = 0; d < kPrefetch; ++d)
            if (d < n_cg) load_group(d, va[d], vb[d]);
        int slot = 0; uint32_t ph = 0;
        for (int cg0 = 0; cg0 < n_cg; cg0 += kPrefetch) {
#pragma unroll
            for (int d = 0; d < kPrefetch; ++d) {
                const int cg = cg0 + d;
                if (cg < n_cg) {
                    mbar_wait(&empty[slot], ph ^ 1u);
                    unsigned char* st = smem + (size_t)slot * p.stage_bytes;
                    // halo: fp32 -> (hi, lo) bf16 planes
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        if (src_off[k] != -2 && !(p.ablate & 1)) {
                            uint4 hi, lo;
                            split8(va[d][k], vb[d][k], hi, lo);
                            *reinterpret_cast<uint4*>(st + dst_off[k]) = hi;
                            *reinterpret_cast<uint4*>(st + kSplitAPlane + dst_off[k]) = lo;
                        }
                    }
                    fence_proxy_async();             // st.shared operands -> tensor core (async proxy) reads
                    mbar_arrive(&full[slot]);
                    if (cg + kPrefetch < n_cg) load_group(cg + kPrefetch, va[d], vb[d]);
                    if (++slot == p.stages) { slot = 0; ph ^= 1u; }
                }
            }
        }
    } else if (warp == 8) {
        // ================= MMA issue =================
        // hi_x against all weight parts (N = parts * nt), lo_x against all but the last (N = (parts - 1) * nt)
        const uint32_t idesc_hi = instr_desc_bf16(p.wparts * nt, false, false);
        const uint32_t idesc_lo = instr_desc_bf16((p.wparts - 1) * nt, false, false);
        int slot = 0; uint32_t ph = 0;
        uint32_t accum = 0;
        for (int cg = 0; cg < n_cg; ++cg) {
            mbar_wait(&full[slot], ph);
            tc_fence_after();
            if (elect_one()) {       // single-thread region behind ONE elect.sync (umma.cuh)
                const uint32_t sA = smem_u32(smem + (size_t)slot * p.stage_bytes);
                const uint64_t a_hi = smem_desc(sA, (uint32_t)kHaloPitch * 16u, kHaloW * 16u);
                const uint64_t a_lo = smem_desc(sA + kSplitAPlane, (uint32_t)kHaloPitch * 16u, kHaloW * 16u);
                // B of a tap: [chunk][part][nt] x 16 bytes: K chunks parts * nt rows apart, 8-row groups 128 B apart
                const uint64_t b0 = smem_desc(sA + 2 * kSplitAPlane, (uint32_t)(p.wparts * nt) * 16u, 128u);
                const uint32_t b_tap = (uint32_t)(2 * p.wparts * nt);          // 16-byte units per tap
#pragma unroll
                for (int tap = 0; tap < ((p.ablate & 4) ? 0 : 9); ++tap) {
                    const uint64_t off = (uint64_t)((tap / 3) * kHaloW + (tap % 3));
                    const uint64_t b = b0 + (uint64_t)tap * b_tap;
                    mma_bf16(tmem_base, a_hi + off, b, idesc_hi, accum);
                    mma_bf16(tmem_base, a_lo + off, b, idesc_lo, 1u);
                    accum = 1;
                }
                mma_commit(&empty[slot]);
                if (cg + 1 == n_cg) mma_commit(acc_full);
            }
            __syncwarp();
            if (++slot == p.stages) { slot = 0; ph ^= 1u; }
        }
    } else {
        // ================= weight copies (while the reduction runs), then the epilogue =================
        {
            const int w_rows = 18 * p.wparts;        // [tap][chunk][part] rows of Cout entries per channel group
            const int w_items = w_rows * nt;         // 16-byte chunks of them in this slice (nt per row, starting at n0)
            // item i = q * nt + n -> source row q, column n; (q, n) advance by divmod(128, nt) per step: no division in the loop
            const int q0 = tid / nt, c0 = tid - q0 * nt, dq = 128 / nt, dn = 128 - dq * nt;
            int slot = 0; uint32_t ph = 0;
            for (int cg = 0; cg < n_cg; ++cg) {
                mbar_wait(&empty[slot], ph ^ 1u);
                const uint4* wsrc = p.wpack + (size_t)cg * w_rows * p.Cout + n0;
                const uint32_t sB = smem_u32(smem + (size_t)slot * p.stage_bytes + 2 * kSplitAPlane);
                int q = q0, n = c0;
                for (int i = tid; i < ((p.ablate & 2) ? 0 : w_items); i += 128) {
                    cp_async16_full(sB + (uint32_t)i * 16u, wsrc + (size_t)q * p.Cout + n);
                    q += dq; n += dn;
                    if (n >= nt) { n -= nt; ++q; }
                }
                cp_async_arrive(&full[slot]);
                if (++slot == p.stages) { slot = 0; ph ^= 1u; }
            }
        }
        const int m = warp * 32 + lane;              // TMEM lane == pixel of the tile
        const int oy = tyi * kTileH + (m >> 3), ox = txi * kTileW + (m & 7);
        const bool valid = oy < p.H && ox < p.W;
        const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
        const int units = nt >> 4;
        mbar_wait(acc_full, 0);
        tc_fence_after();
        const int n_terms = p.wparts;
        auto load_sum16 = [&](int u, float (&v)[16]) {       // sum over the column groups of the weight parts, 16 channels
            tmem_ld16(taddr + u * 16, v);
            tmem_wait_ld();
            for (int t = 1; t < n_terms; ++t) {
                float w16[16];
                tmem_ld16(taddr + t * nt + u * 16, w16);
                tmem_wait_ld();
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] += w16[j];
            }
        };
        float scale = 1.0f;
        if (kPN) {
            float ss = 0.0f;
            for (int u = 0; u < units; ++u) {
                float v[16];
                load_sum16(u, v);
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    float t = v[j];
                    if (kBA) { t += p.bias ? __ldg(p.bias + n0 + u * 16 + j) : 0.0f; t = p.lrelu ? fmaxf(t, 0.2f * t) : t; }
                    ss = fmaf(t, t, ss);
                }
            }
            scale = 1.0f / sqrtf(ss / (float)nt + 1e-8f);
            if (valid && p.inv_norm) p.inv_norm[((size_t)b * p.H + oy) * p.W + ox] = scale;
        }
        const size_t pix = ((size_t)b * p.H + oy) * p.W + ox;
        float* dst = p.y + pix * p.Cout + n0;
        __nv_bfloat16* dst16 = p.y16 ? p.y16 + pix * p.Cout + n0 : nullptr;
        for (int u = 0; u < units; ++u) {
            float v[16];
            load_sum16(u, v);
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                float t = v[j];
                if (kBA) { t += p.bias ? __ldg(p.bias + n0 + u * 16 + j) : 0.0f; t = p.lrelu ? fmaxf(t, 0.2f * t) : t; }
                if (kPN) t *= scale;
                v[j] = t;
            }
            if (valid && !(p.ablate & 8)) {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    *reinterpret_cast<float4*>(dst + u * 16 + 4 * j) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                if (dst16) {
                    uint32_t pk[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
                        pk[j] = *reinterpret_cast<const uint32_t*>(&h);
                    }
                    *reinterpret_cast<uint4*>(dst16 + u * 16) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                    *reinterpret_cast<uint4*>(dst16 + u * 16 + 8) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 8) tmem_dealloc(tmem_base, p.tmem_cols);
}

__global__ void k_pack_weights_split(const float* __restrict__ w, int Cout, int Cin, int transpose_flip, int wparts,
                                     __nv_bfloat16* __restrict__ out) {
    pdl_trigger();
    pdl_wait();
    pack_weights_split_range(w, Cout, Cin, transpose_flip, wparts, out, blockIdx.x * blockDim.x + threadIdx.x, gridDim.x * blockDim.x);
}

struct SplitPlan { int Nt, n_slices, stages, tmem_cols; unsigned stage_bytes; size_t smem; };

static SplitPlan plan_split(int Cin, int Cout, bool full_n, int n_tiles, int wparts) {
    SplitPlan pl{};
    // layers with at least 128 tiles run TWO CTAs per SM (each gets half of the shared memory, i.e. a shorter ring and
    // narrower slices): the kernel is not persistent, so the prologue / load / MMA / epilogue phases of neighbouring tiles
    // only overlap across CTAs (scripts/sweep_split.py: 64 x 64 layers at batch 16 -20 %, 32 x 32 at batch 16 -18 %; layers
    // with 64 tiles or fewer lose 20 % and keep one CTA with a deep ring).  MG_SPLIT_2CTA_TILES overrides (0 = never).
    const int two_cta_tiles = getenv("MG_SPLIT_2CTA_TILES") ? atoi(getenv("MG_SPLIT_2CTA_TILES")) : 128;
    const size_t budget = (!full_n && two_cta_tiles > 0 && n_tiles >= two_cta_tiles ? 108 : 216) * 1024;
    // slice width.  An M128 x N x K16 MMA from shared memory costs max(32, N / 2) cycles (the A tile is re-read by every
    // MMA), so N <= 64 is as fast per CTA as it gets: layers with fewer tiles than SMs (latency bound) use narrow slices
    // and a deep ring; layers with many tiles use the widest slice that still leaves two stages (fewer CTAs re-reading
    // the same halo)
    // (one MMA spans parts * Nt <= 256 columns: Nt <= 80 with three parts, <= 128 with two)
    const int max_nt = n_tiles >= 148 ? (wparts == 3 ? 80 : 96) : (wparts == 3 ? 48 : 64);
    int slices = full_n ? 1 : (Cout + max_nt - 1) / max_nt;
    // few tiles (tiny images): spread the weight traffic over more CTAs
    while (!full_n && n_tiles * slices < 96 && (Cout / 16 + slices) / (slices + 1) >= 2) ++slices;
    int Nt = ((Cout / 16 + slices - 1) / slices) * 16;
    for (;; Nt -= 16) {
        if (Nt < 16) return pl;
        const unsigned stage = (unsigned)align_up((size_t)2 * kSplitAPlane + (size_t)288 * wparts * Nt, 128);
        int stages = (int)((budget - 256) / stage);
        if (stages < 2) { if (full_n) { if (stages < 1) return pl; } else continue; }
        if (stages > kSplitStagesMax) stages = kSplitStagesMax;
        const int n_cg = Cin / 16;
        if (stages > n_cg) stages = n_cg;
        pl.Nt = Nt; pl.n_slices = (Cout + Nt - 1) / Nt; pl.stages = stages; pl.stage_bytes = stage;
        pl.smem = (size_t)stages * stage + 256;
        int cols = 32; while (cols < wparts * Nt) cols <<= 1;      // `parts` column groups of Nt accumulator columns
        pl.tmem_cols = cols;
        return pl;
    }
}

}  // namespace mg

using namespace mg;

extern "C" {

size_t mg_conv3x3_split_workspace_bytes(int Cin, int Cout) {
    return align_up((size_t)9 * Cin * Cout * 6, 256);       // room for three parts
}

// mode: bit 0 data-gradient orientation, bit 1 three parts (hi + mid + lo, for mg_conv3x3_split_f32 with flag 32)
int mg_conv3x3_split_pack_weights(const float* w_f32, int Cin, int Cout, int mode, void* packed, size_t packed_bytes, mgStream stream) {
    if (!w_f32 || !packed) return MG_ERR_BAD_ARG;
    if (Cin < 16 || Cout < 16 || (Cin & 15) || (Cout & 15) || Cin > 256 || Cout > 256) return MG_ERR_UNSUPPORTED;
    const int dgrad = mode & 1, wparts = (mode & 2) ? 3 : 2;
    if (packed_bytes < (size_t)9 * Cin * Cout * 2 * wparts) return MG_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    ProfScope ps("k_pack_weights_split", st);
    // Cin, Cout are those of the GEMM (for dgrad: Cin = channels of dY = forward Cout)
    const int fwd_cout = dgrad ? Cin : Cout, fwd_cin = dgrad ? Cout : Cin;
    const int total = 9 * Cin * Cout;
    launch_pdl(k_pack_weights_split, dim3((total + 255) / 256), dim3(256), 0, st, w_f32, fwd_cout, fwd_cin, dgrad ? 1 : 0, wparts, (__nv_bfloat16*)packed);
    return check_launch("k_pack_weights_split");
}

// flags as mg_conv3x3_bf16: 1 LeakyReLU(0.2), 2 PixelNorm, 4 nearest x2 upsampled input, 8 data gradient; 32: `packed` holds
// three weight parts (mode bit 1 of the pack call)
int mg_conv3x3_split_f32(const float* x, const void* packed, const float* bias, float* y, void* y_bf16, float* inv_norm,
                         int B, int H, int W, int Cin, int Cout, int flags, mgStream stream) {
    if (!x || !y || !packed) return MG_ERR_BAD_ARG;
    if (B <= 0 || H <= 0 || W <= 0 || Cin < 16 || Cout < 16 || (Cin & 15) || (Cout & 15) || Cin > 256 || Cout > 256) return MG_ERR_UNSUPPORTED;
    const bool ups = (flags & 4) != 0, pn = (flags & 2) != 0;
    if (ups && ((H | W) & 1)) return MG_ERR_BAD_ARG;
    SplitParams p{};
    p.tiles_x = (W + kTileW - 1) / kTileW; p.tiles_y = (H + kTileH - 1) / kTileH;
    const long long n_tiles = (long long)B * p.tiles_x * p.tiles_y;
    if (n_tiles >= (1ll << 30)) return MG_ERR_UNSUPPORTED;
    const int wparts = (flags & 32) ? 3 : 2;
    SplitPlan pl = plan_split(Cin, Cout, pn, (int)n_tiles, wparts);
    if (pl.Nt == 0 || (pn && pl.n_slices != 1)) return MG_ERR_UNSUPPORTED;
    p.x = x; p.wpack = (const uint4*)packed; p.bias = bias; p.y = y; p.y16 = (__nv_bfloat16*)y_bf16; p.inv_norm = inv_norm;
    p.B = B; p.H = H; p.W = W; p.Hin = ups ? H / 2 : H; p.Win = ups ? W / 2 : W; p.Cin = Cin; p.Cout = Cout;
    p.upsample = ups; p.lrelu = flags & 1; p.pixelnorm = pn;
    p.Nt = pl.Nt; p.stages = pl.stages; p.tmem_cols = pl.tmem_cols; p.stage_bytes = pl.stage_bytes; p.wparts = wparts;
    p.ablate = getenv("MG_SPLIT_ABLATE") ? atoi(getenv("MG_SPLIT_ABLATE")) : 0;
    const bool ba = bias != nullptr || (flags & 1);
    auto kern = pn ? (ba ? k_conv3x3_split<true, true> : k_conv3x3_split<true, false>)
                   : (ba ? k_conv3x3_split<false, true> : k_conv3x3_split<false, false>);
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem);
    cudaStream_t st = (cudaStream_t)stream;
    {
        ProfScope ps((flags & 8) ? "k_conv3x3_split_dgrad" : "k_conv3x3_split_fprop", st);
        launch_pdl(kern, dim3((unsigned)n_tiles, pl.n_slices), dim3(kSplitThreads), pl.smem, st, p);
    }
    return check_launch("k_conv3x3_split");
}

}  // extern "C"
