// Memory-bound companions of the convolution kernels (sm_100a, CUDA cores, 16-byte vector accesses):
//
//   rgb_expand    2 -> C   1x1 convolution on fp32 NCHW planes -> bf16 NHWC, + bias, + LeakyReLU or a mask
//                          multiply           (MagPhaseLayer, reference networks/discriminator.py:37-50)
//   rgb_project   C -> 2   1x1 convolution on bf16 NHWC -> fp32 NCHW planes, + bias, + tanh, optional mask on the
//                          input            (ToMagnPhaseLayer, networks/generator.py:43-52; and the data gradient
//                          of rgb_expand)
//   rgb_wgrad     sum over pixels of (masked) bf16 NHWC tensor (x) fp32 planes -> [C][2] (+ [C]) weight / bias grads
//   pool2 / unpool2   2x2 average pooling on bf16 NHWC and its adjoint (AvgPool2d(2,2), discriminator.py:24,131)
//
// "mask" = LeakyReLU(0.2) derivative recovered from the sign of a saved activation h: 1 where h > 0 else 0.2.
// The three rgb_* maps are closed under differentiation (each one's derivative w.r.t. either argument is another
// of the three), which is what the gradient penalty's double backward needs.
#include "common.cuh"
#include <cuda_bf16.h>

namespace mg {
// Per-channel sums over pixels, in two deterministic steps and without atomics.  (Every block ending with C global
// atomics on the same cache line serialises in one L2 slice: ~27 cycles per warp request, 10-20 us for a few hundred
// blocks -- more than the streaming time of most layers.)
//   step 1 (inside the streaming kernel): a thread owns one 8-channel chunk for all its pixels (the grid stride is a
//           multiple of C/8); the block transposes its 256 x 8 register partials through shared memory and writes ONE
//           row part[blockIdx.x][C];
//   step 2 (k_colsum_reduce): gb[c] = sum over rows, fixed order, gb overwritten (no zero fill needed).
__device__ __forceinline__ void block_colsum_to_row(const float (&acc)[8], float* s_t, float* __restrict__ row,
                                                    int64_t first, int C8) {
    // s_t [256][9]: thread-major, padded
#pragma unroll
    for (int j = 0; j < 8; ++j) s_t[threadIdx.x * 9 + j] = acc[j];
    __syncthreads();
    const int C = C8 * 8;
    const int q0 = (int)((first - threadIdx.x) % C8);      // chunk of thread 0 of this block
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const int q = c >> 3, j = c & 7;
        int t = q - q0; if (t < 0) t += C8;                 // first thread of the block that owns chunk q
        float v = 0.0f;
        for (; t < (int)blockDim.x; t += C8) v += s_t[t * 9 + j];
        row[c] = v;
    }
}

// gb[c] = sum_r part[r][c]; block = 8 warps x 32 consecutive channels, warp w takes rows w, w+8, ...
__global__ void __launch_bounds__(256)
k_colsum_reduce(const float* __restrict__ part, float* __restrict__ gb, int C, int rows) {
    pdl_trigger();
    pdl_wait();
    __shared__ float red[8][33];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + lane;
    float s0 = 0.0f, s1 = 0.0f;
    if (c < C) {
        int r = w;
        for (; r + 8 < rows; r += 16) { s0 += part[(size_t)r * C + c]; s1 += part[(size_t)(r + 8) * C + c]; }
        if (r < rows) s0 += part[(size_t)r * C + c];
    }
    red[w][lane] = s0 + s1;
    __syncthreads();
    if (w == 0 && c < C) {
        float v = 0.0f;
#pragma unroll
        for (int k = 0; k < 8; ++k) v += red[k][lane];
        gb[c] = v;
    }
}

static const int kColsumMaxBlocks = 148 * 8 + 32;

// grid of a column-sum kernel: blocks * 256 a multiple of C8 so that a thread always meets the same 8-channel chunk
static unsigned colsum_blocks(int64_t total, int C8) {
    int64_t blocks = (total + 1023) / 1024;
    if (blocks < 1) blocks = 1;
    if (blocks > 148 * 8) blocks = 148 * 8;
    while ((blocks * 256) % C8) ++blocks;
    return (unsigned)blocks;
}

__device__ __forceinline__ float lrelu02(float v) { return v > 0.0f ? v : 0.2f * v; }

union Pack8 { uint4 u; __nv_bfloat162 h[4]; };

// ---- 2 -> C ------------------------------------------------------------------------------------
// x [B][2][HW] fp32, w [C][2], b [C] or null, mask_src [B][HW][C] bf16 or null, y [B][HW][C] bf16
// mode: 0 none, 1 LeakyReLU(0.2) on the result, 2 multiply the result by mask(mask_src)
__global__ void __launch_bounds__(256)
k_rgb_expand(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b,
             const __nv_bfloat16* __restrict__ mask_src, __nv_bfloat16* __restrict__ y, int64_t HW, int C, int mode, int64_t total) {
    pdl_trigger();
    pdl_wait();
    extern __shared__ float sw[];           // [C][2] then [C] bias
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) sw[i] = w[i];
    for (int i = threadIdx.x; i < C; i += blockDim.x) sw[2 * C + i] = b ? b[i] : 0.0f;
    __syncthreads();
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < total; p += (int64_t)gridDim.x * blockDim.x) {
        const int64_t bi = p / HW, q = p - bi * HW;
        const float x0 = x[(bi * 2) * HW + q], x1 = x[(bi * 2 + 1) * HW + q];
        uint4* dst = reinterpret_cast<uint4*>(y + p * C);
        const uint4* msk = mask_src ? reinterpret_cast<const uint4*>(mask_src + p * C) : nullptr;
        for (int c0 = 0; c0 < C; c0 += 8) {
            Pack8 o, m;
            if (mode == 2) m.u = msk[c0 >> 3];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int c = c0 + 2 * j;
                float v0 = fmaf(sw[2 * c], x0, fmaf(sw[2 * c + 1], x1, sw[2 * C + c]));
                float v1 = fmaf(sw[2 * c + 2], x0, fmaf(sw[2 * c + 3], x1, sw[2 * C + c + 1]));
                if (mode == 1) { v0 = lrelu02(v0); v1 = lrelu02(v1); }
                if (mode == 2) {
                    const float2 mm = __bfloat1622float2(m.h[j]);
                    v0 *= mm.x > 0.0f ? 1.0f : 0.2f; v1 *= mm.y > 0.0f ? 1.0f : 0.2f;
                }
                o.h[j] = __floats2bfloat162_rn(v0, v1);
            }
            dst[c0 >> 3] = o.u;
        }
    }
}

// ---- C -> 2 ------------------------------------------------------------------------------------
// a [B][HW][C] bf16, w2 = two rows of C weights: row k at w2[k * row_stride + c * col_stride] (so both a [2][C]
// forward weight and the transpose of a [C][2] weight can be passed), bias [2] or null, mask_src or null,
// out [B][2][HW] fp32.  act: 0 none, 1 tanh
__global__ void __launch_bounds__(256)
k_rgb_project(const __nv_bfloat16* __restrict__ a, const float* __restrict__ w2, int row_stride, int col_stride,
              const float* __restrict__ bias, const __nv_bfloat16* __restrict__ mask_src, float* __restrict__ out,
              int64_t HW, int C, int act, int64_t total) {
    pdl_trigger();
    pdl_wait();
    extern __shared__ float sw[];           // [2][C]
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) { const int k = i / C, c = i - k * C; sw[i] = w2[k * row_stride + c * col_stride]; }
    __syncthreads();
    const float b0 = bias ? bias[0] : 0.0f, b1 = bias ? bias[1] : 0.0f;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < total; p += (int64_t)gridDim.x * blockDim.x) {
        const int64_t bi = p / HW, q = p - bi * HW;
        const uint4* src = reinterpret_cast<const uint4*>(a + p * C);
        const uint4* msk = mask_src ? reinterpret_cast<const uint4*>(mask_src + p * C) : nullptr;
        float s0 = b0, s1 = b1;
        for (int c0 = 0; c0 < C; c0 += 8) {
            Pack8 v, m;
            v.u = __ldg(src + (c0 >> 3));
            if (msk) m.u = __ldg(msk + (c0 >> 3));
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float2 f = __bfloat1622float2(v.h[j]);
                if (msk) { const float2 mm = __bfloat1622float2(m.h[j]); f.x *= mm.x > 0.0f ? 1.0f : 0.2f; f.y *= mm.y > 0.0f ? 1.0f : 0.2f; }
                const int c = c0 + 2 * j;
                s0 = fmaf(f.x, sw[c], s0); s0 = fmaf(f.y, sw[c + 1], s0);
                s1 = fmaf(f.x, sw[C + c], s1); s1 = fmaf(f.y, sw[C + c + 1], s1);
            }
        }
        if (act == 1) { s0 = tanhf(s0); s1 = tanhf(s1); }
        out[(bi * 2) * HW + q] = s0;
        out[(bi * 2 + 1) * HW + q] = s1;
    }
}

// ---- weight / bias gradient of the 1x1 layers ---------------------------------------------------
// gw[c][0..1] += sum_p g[p][c] * mask * x[b][0..1][q],  gb[c] += sum_p g[p][c] * mask
// grid (blocks, C/8): each thread owns 8 channels of a strided set of pixels; warp shuffle + atomics
__global__ void __launch_bounds__(256)
k_rgb_wgrad(const __nv_bfloat16* __restrict__ g, const __nv_bfloat16* __restrict__ mask_src, const float* __restrict__ x,
            float* __restrict__ gw, float* __restrict__ gb, int64_t HW, int C, int64_t total) {
    pdl_trigger();
    pdl_wait();
    const int c0 = blockIdx.y * 8;
    float acc[8][3];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j][0] = acc[j][1] = acc[j][2] = 0.0f;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < total; p += (int64_t)gridDim.x * blockDim.x) {
        const int64_t bi = p / HW, q = p - bi * HW;
        const float x0 = x[(bi * 2) * HW + q], x1 = x[(bi * 2 + 1) * HW + q];
        Pack8 v, m;
        v.u = __ldg(reinterpret_cast<const uint4*>(g + p * C + c0));
        if (mask_src) m.u = __ldg(reinterpret_cast<const uint4*>(mask_src + p * C + c0));
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float2 f = __bfloat1622float2(v.h[j]);
            if (mask_src) { const float2 mm = __bfloat1622float2(m.h[j]); f.x *= mm.x > 0.0f ? 1.0f : 0.2f; f.y *= mm.y > 0.0f ? 1.0f : 0.2f; }
            acc[2 * j][0] = fmaf(f.x, x0, acc[2 * j][0]); acc[2 * j][1] = fmaf(f.x, x1, acc[2 * j][1]); acc[2 * j][2] += f.x;
            acc[2 * j + 1][0] = fmaf(f.y, x0, acc[2 * j + 1][0]); acc[2 * j + 1][1] = fmaf(f.y, x1, acc[2 * j + 1][1]); acc[2 * j + 1][2] += f.y;
        }
    }
    __shared__ float red[8][24];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            float v = acc[j][k];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0) red[warp][j * 3 + k] = v;
        }
    __syncthreads();
    if (threadIdx.x < 24) {
        float v = 0.0f;
        for (int w = 0; w < 8; ++w) v += red[w][threadIdx.x];
        const int j = threadIdx.x / 3, k = threadIdx.x % 3;
        if (k < 2) atomicAdd(&gw[(c0 + j) * 2 + k], v);
        else if (gb) atomicAdd(&gb[c0 + j], v);
    }
}

// ---- 2x2 average pooling, bf16 NHWC ---------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_pool2(const __nv_bfloat16* __restrict__ in, __nv_bfloat16* __restrict__ out, int Ho, int Wo, int C8, int64_t total, float scale) {
    pdl_trigger();
    pdl_wait();
    // one thread = 8 channels of one output pixel
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C8);
        int64_t r = i / C8;
        const int ox = (int)(r % Wo); r /= Wo;
        const int oy = (int)(r % Ho);
        const int64_t b = r / Ho;
        const uint4* src = reinterpret_cast<const uint4*>(in) + ((b * 2 * Ho + 2 * oy) * (int64_t)(2 * Wo) + 2 * ox) * C8 + c;
        Pack8 v00, v01, v10, v11, o;
        v00.u = __ldg(src); v01.u = __ldg(src + C8);
        v10.u = __ldg(src + (int64_t)2 * Wo * C8); v11.u = __ldg(src + (int64_t)2 * Wo * C8 + C8);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float2 a = __bfloat1622float2(v00.h[j]), bb = __bfloat1622float2(v01.h[j]);
            const float2 cc = __bfloat1622float2(v10.h[j]), d = __bfloat1622float2(v11.h[j]);
            o.h[j] = __floats2bfloat162_rn(scale * ((a.x + bb.x) + (cc.x + d.x)), scale * ((a.y + bb.y) + (cc.y + d.y)));
        }
        reinterpret_cast<uint4*>(out)[i] = o.u;
    }
}

// adjoint: out[b][2y+i][2x+j][c] = 0.25 * in[b][y][x][c]
__global__ void __launch_bounds__(256)
k_unpool2(const __nv_bfloat16* __restrict__ in, __nv_bfloat16* __restrict__ out, int Hi, int Wi, int C8, int64_t total) {
    pdl_trigger();
    pdl_wait();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C8);
        int64_t r = i / C8;
        const int x = (int)(r % Wi); r /= Wi;
        const int y = (int)(r % Hi);
        const int64_t b = r / Hi;
        Pack8 v, o;
        v.u = __ldg(reinterpret_cast<const uint4*>(in) + i);
#pragma unroll
        for (int j = 0; j < 4; ++j) { const float2 f = __bfloat1622float2(v.h[j]); o.h[j] = __floats2bfloat162_rn(0.25f * f.x, 0.25f * f.y); }
        uint4* dst = reinterpret_cast<uint4*>(out) + ((b * 2 * Hi + 2 * y) * (int64_t)(2 * Wi) + 2 * x) * C8 + c;
        dst[0] = o.u; dst[C8] = o.u; dst[(int64_t)2 * Wi * C8] = o.u; dst[(int64_t)2 * Wi * C8 + C8] = o.u;
    }
}

// ---- AvgPool2d(2,2) on fp32 NCHW planes (the network input on the fade-in path, discriminator.py:130-133) ----------
// adjoint 0: in [n][2Ho][2Wo] -> out [n][Ho][Wo] = ((a + b) + c) + d) * 0.25 (torch's summation order: bit-identical);
// adjoint 1: in [n][Ho][Wo] -> out [n][2Ho][2Wo] = 0.25 * in replicated (its backward).  One thread per low-res element.
__global__ void __launch_bounds__(256)
k_pool2_planes(const float* __restrict__ in, float* __restrict__ out, int Ho, int Wo, int64_t total, int adjoint) {
    pdl_trigger();
    pdl_wait();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int x = (int)(i % Wo);
        const int64_t r = i / Wo;                       // plane * Ho + y
        const int64_t hi = (r * 2) * (int64_t)(2 * Wo) + 2 * x;        // top-left element of the 2x2 block
        if (!adjoint) {
            const float2 t = *reinterpret_cast<const float2*>(in + hi);
            const float2 b = *reinterpret_cast<const float2*>(in + hi + 2 * Wo);
            out[i] = __fmul_rn(__fadd_rn(__fadd_rn(__fadd_rn(t.x, t.y), b.x), b.y), 0.25f);
        } else {
            const float v = 0.25f * in[i];
            *reinterpret_cast<float2*>(out + hi) = make_float2(v, v);
            *reinterpret_cast<float2*>(out + hi + 2 * Wo) = make_float2(v, v);
        }
    }
}

// ---- LeakyReLU backward + bias gradient ------------------------------------------------------------
// gz = gy * mask(y) (bf16 NHWC), gb[c] += sum_pixels gz[.,c].  grid.x blocks x 256 threads; a thread keeps the same
// 8-channel chunk for all its pixels (grid stride is a multiple of C/8), so its 8 partial sums stay in registers.
__global__ void __launch_bounds__(256)
k_lrelu_bwd(const __nv_bfloat16* __restrict__ gy, const __nv_bfloat16* __restrict__ y, __nv_bfloat16* __restrict__ gz,
            float* __restrict__ part, int C8, int64_t total, int64_t stride) {
    pdl_trigger();
    pdl_wait();
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.0f;
    const int64_t first = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    // 4 independent 16-byte loads per operand in flight per thread (memory-level parallelism)
    int64_t i = first;
    for (; i + 3 * stride < total; i += 4 * stride) {
        Pack8 g[4], m[4], o[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            g[u].u = __ldcs(reinterpret_cast<const uint4*>(gy) + i + u * stride);
            m[u].u = __ldcs(reinterpret_cast<const uint4*>(y) + i + u * stride);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float2 f = __bfloat1622float2(g[u].h[j]);
                const float2 mm = __bfloat1622float2(m[u].h[j]);
                f.x *= mm.x > 0.0f ? 1.0f : 0.2f; f.y *= mm.y > 0.0f ? 1.0f : 0.2f;
                o[u].h[j] = __floats2bfloat162_rn(f.x, f.y);
                acc[2 * j] += f.x; acc[2 * j + 1] += f.y;
            }
            reinterpret_cast<uint4*>(gz)[i + u * stride] = o[u].u;
        }
    }
    for (; i < total; i += stride) {
        Pack8 g, m, o;
        g.u = __ldg(reinterpret_cast<const uint4*>(gy) + i);
        m.u = __ldg(reinterpret_cast<const uint4*>(y) + i);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float2 f = __bfloat1622float2(g.h[j]);
            const float2 mm = __bfloat1622float2(m.h[j]);
            f.x *= mm.x > 0.0f ? 1.0f : 0.2f; f.y *= mm.y > 0.0f ? 1.0f : 0.2f;
            o.h[j] = __floats2bfloat162_rn(f.x, f.y);
            acc[2 * j] += f.x; acc[2 * j + 1] += f.y;
        }
        reinterpret_cast<uint4*>(gz)[i] = o.u;
    }
    if (part) {     // this block's row of per-channel sums (k_colsum_reduce adds the rows)
        __shared__ float s_t[256 * 9];
        block_colsum_to_row(acc, s_t, part + (size_t)blockIdx.x * C8 * 8, first, C8);
    }
}

// ---- avg-pool adjoint fused with the LeakyReLU backward of the tensor that was pooled ------------------------------
// (ConvBlock of the discriminator: conv -> LeakyReLU -> AvgPool2d(2,2), discriminator.py:15-24.)
//   gz[b][2y+i][2x+j][c] = 0.25 * gp[b][y][x][c] * mask(h[b][2y+i][2x+j][c]),  gb[c] = sum over pixels of gz
// One pass (read gp once and h once, write gz once) instead of k_unpool2 (write the un-pooled gradient) followed by
// k_lrelu_bwd (read it back together with h): 2.25 instead of 4.25 tensor volumes.  Bit-identical values: 0.25 * g is
// exact in bf16.  Threads walk the LOW-resolution chunks with the same thread <-> 8-channel-chunk ownership as
// k_lrelu_bwd.
__global__ void __launch_bounds__(256)
k_unpool2_lrelu_bwd(const __nv_bfloat16* __restrict__ gp, const __nv_bfloat16* __restrict__ h, __nv_bfloat16* __restrict__ gz,
                    float* __restrict__ part, int Hi, int Wi, int C8, int64_t total, int64_t stride) {
    pdl_trigger();
    pdl_wait();
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.0f;
    const int64_t first = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t row = (int64_t)2 * Wi * C8;                 // one full-resolution image row, in 16-byte chunks
    for (int64_t i = first; i < total; i += stride) {
        const int c = (int)(i % C8);
        int64_t r = i / C8;
        const int x = (int)(r % Wi); r /= Wi;
        const int y = (int)(r % Hi);
        const int64_t b = r / Hi;
        const int64_t o0 = ((b * 2 * Hi + 2 * y) * (int64_t)(2 * Wi) + 2 * x) * C8 + c;
        Pack8 g, m[4], o[4];
        g.u = __ldg(reinterpret_cast<const uint4*>(gp) + i);
        const uint4* hp = reinterpret_cast<const uint4*>(h) + o0;
        m[0].u = __ldcs(hp); m[1].u = __ldcs(hp + C8); m[2].u = __ldcs(hp + row); m[3].u = __ldcs(hp + row + C8);
        float2 gq[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) { const float2 f = __bfloat1622float2(g.h[j]); gq[j] = make_float2(0.25f * f.x, 0.25f * f.y); }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 mm = __bfloat1622float2(m[q].h[j]);
                const float vx = gq[j].x * (mm.x > 0.0f ? 1.0f : 0.2f), vy = gq[j].y * (mm.y > 0.0f ? 1.0f : 0.2f);
                o[q].h[j] = __floats2bfloat162_rn(vx, vy);
                acc[2 * j] += vx; acc[2 * j + 1] += vy;
            }
        }
        uint4* dst = reinterpret_cast<uint4*>(gz) + o0;
        dst[0] = o[0].u; dst[C8] = o[1].u; dst[row] = o[2].u; dst[row + C8] = o[3].u;
    }
    if (part) {
        __shared__ float s_t[256 * 9];
        block_colsum_to_row(acc, s_t, part + (size_t)blockIdx.x * C8 * 8, first, C8);
    }
}

// ---- PixelNorm + LeakyReLU backward (generator half-block, layers.py:11-17 + generator.py:23,38) -----------------
// o = t / n with t = lrelu(z), n = sqrt(mean_c t^2 + eps), inv = 1/n saved by the forward kernel.
//   g_t = (g_o - o * mean_c(g_o * o)) * inv ;  g_z = g_t * (o > 0 ? 1 : 0.2) ;  gb[c] += sum_pixels g_z
// one thread per pixel (channels contiguous in NHWC), two passes over its C channels
__global__ void __launch_bounds__(256)
k_pixelnorm_lrelu_bwd(const __nv_bfloat16* __restrict__ go, const __nv_bfloat16* __restrict__ o, const float* __restrict__ inv,
                      __nv_bfloat16* __restrict__ gz, int C, int64_t n_pixels) {
    pdl_trigger();
    pdl_wait();
    const int C8 = C >> 3;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n_pixels; p += (int64_t)gridDim.x * blockDim.x) {
        const uint4* g4 = reinterpret_cast<const uint4*>(go + p * C);
        const uint4* o4 = reinterpret_cast<const uint4*>(o + p * C);
        float dot = 0.0f;
        for (int c = 0; c < C8; ++c) {
            Pack8 a, b; a.u = __ldg(g4 + c); b.u = __ldg(o4 + c);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 x = __bfloat1622float2(a.h[j]), y = __bfloat1622float2(b.h[j]);
                dot = fmaf(x.x, y.x, dot); dot = fmaf(x.y, y.y, dot);
            }
        }
        dot /= (float)C;
        const float s = inv[p];
        uint4* z4 = reinterpret_cast<uint4*>(gz + p * C);
        for (int c = 0; c < C8; ++c) {
            Pack8 a, b, r; a.u = __ldg(g4 + c); b.u = __ldg(o4 + c);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 x = __bfloat1622float2(a.h[j]), y = __bfloat1622float2(b.h[j]);
                float t0 = (x.x - y.x * dot) * s, t1 = (x.y - y.y * dot) * s;
                t0 *= y.x > 0.0f ? 1.0f : 0.2f; t1 *= y.y > 0.0f ? 1.0f : 0.2f;
                r.h[j] = __floats2bfloat162_rn(t0, t1);
            }
            z4[c] = r.u;
        }
    }
}

// part[blockIdx.x][c] = this block's sum over its pixels of g[.,c]  (bf16 NHWC); same thread <-> channel-chunk ownership
// as k_lrelu_bwd
__global__ void __launch_bounds__(256)
k_colsum(const __nv_bfloat16* __restrict__ g, float* __restrict__ part, int C8, int64_t total, int64_t stride) {
    pdl_trigger();
    pdl_wait();
    __shared__ float s_t[256 * 9];
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.0f;
    const int64_t first = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (int64_t i = first; i < total; i += stride) {
        Pack8 v; v.u = __ldg(reinterpret_cast<const uint4*>(g) + i);
#pragma unroll
        for (int j = 0; j < 4; ++j) { const float2 f = __bfloat1622float2(v.h[j]); acc[2 * j] += f.x; acc[2 * j + 1] += f.y; }
    }
    block_colsum_to_row(acc, s_t, part + (size_t)blockIdx.x * C8 * 8, first, C8);
}

static unsigned grid_for(int64_t items, int per_block = 256, int cap = 148 * 16) {
    int64_t g = (items + per_block - 1) / per_block;
    return (unsigned)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace mg

using namespace mg;

extern "C" {

int mg_rgb_expand_bf16(const float* x, const float* w, const float* b, const void* mask_src, void* y,
                       int B, int64_t HW, int C, int mode, mgStream stream) {
    if (!x || !w || !y || B <= 0 || HW <= 0 || C < 8 || (C & 7) || C > 512) return MG_ERR_BAD_ARG;
    if (mode == 2 && !mask_src) return MG_ERR_BAD_ARG;
    const int64_t total = (int64_t)B * HW;
    cudaStream_t st = (cudaStream_t)stream;
    ProfScope ps("k_rgb_expand", st);
    launch_pdl(k_rgb_expand, dim3(grid_for(total)), dim3(256), 3 * C * sizeof(float), st, x, w, b, (const __nv_bfloat16*)mask_src, (__nv_bfloat16*)y, HW, C, mode, total);
    return check_launch("k_rgb_expand");
}

int mg_rgb_project_bf16(const void* a, const float* w2, int row_stride, int col_stride, const float* bias, const void* mask_src,
                        float* out, int B, int64_t HW, int C, int act, mgStream stream) {
    if (!a || !w2 || !out || B <= 0 || HW <= 0 || C < 8 || (C & 7) || C > 512) return MG_ERR_BAD_ARG;
    const int64_t total = (int64_t)B * HW;
    cudaStream_t st = (cudaStream_t)stream;
    ProfScope ps("k_rgb_project", st);
    launch_pdl(k_rgb_project, dim3(grid_for(total)), dim3(256), 2 * C * sizeof(float), st, (const __nv_bfloat16*)a, w2, row_stride, col_stride, bias,
               (const __nv_bfloat16*)mask_src, out, HW, C, act, total);
    return check_launch("k_rgb_project");
}

int mg_rgb_wgrad_bf16(const void* g, const void* mask_src, const float* x, float* gw, float* gb, int B, int64_t HW, int C, mgStream stream) {
    if (!g || !x || !gw || B <= 0 || HW <= 0 || C < 8 || (C & 7) || C > 512) return MG_ERR_BAD_ARG;
    const int64_t total = (int64_t)B * HW;
    cudaStream_t st = (cudaStream_t)stream;
    ProfScope ps("k_rgb_wgrad", st);
    const unsigned gx = grid_for(total, 256 * 8, 148 * 2);      // every block ends with 24 same-address atomics: keep the tail short
    launch_pdl(k_rgb_wgrad, dim3(gx, C / 8), dim3(256), 0, st, (const __nv_bfloat16*)g, (const __nv_bfloat16*)mask_src, x, gw, gb, HW, C, total);
    return check_launch("k_rgb_wgrad");
}

int mg_pool2_planes_f32(const float* in, float* out, int64_t n_planes, int Ho, int Wo, int adjoint, mgStream stream) {
    if (!in || !out || n_planes <= 0 || Ho <= 0 || Wo <= 0) return MG_ERR_BAD_ARG;
    const int64_t total = n_planes * Ho * Wo;
    cudaStream_t st = (cudaStream_t)stream;
    ProfScope ps("k_pool2_planes", st);
    launch_pdl(k_pool2_planes, dim3(grid_for(total)), dim3(256), 0, st, in, out, Ho, Wo, total, adjoint ? 1 : 0);
    return check_launch("k_pool2_planes");
}

size_t mg_colsum_workspace_bytes(int C) {
    return C > 0 ? align_up((size_t)kColsumMaxBlocks * C * sizeof(float), 256) : 0;
}

// gb (optional, OVERWRITTEN) needs `ws` of mg_colsum_workspace_bytes(C)
int mg_lrelu_bwd_bf16(const void* gy, const void* y, void* gz, float* gb, void* ws, size_t ws_bytes, int64_t n_pixels, int C, mgStream stream) {
    if (!gy || !y || !gz || n_pixels <= 0 || C < 8 || (C & 7)) return MG_ERR_BAD_ARG;
    if (gb && (!ws || ws_bytes < mg_colsum_workspace_bytes(C))) return MG_ERR_WORKSPACE;
    const int C8 = C / 8;
    const int64_t total = n_pixels * C8;
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned blocks = colsum_blocks(total, C8);
    {
        ProfScope ps("k_lrelu_bwd", st);
        launch_pdl(k_lrelu_bwd, dim3(blocks), dim3(256), 0, st, (const __nv_bfloat16*)gy, (const __nv_bfloat16*)y, (__nv_bfloat16*)gz,
                   gb ? (float*)ws : (float*)nullptr, C8, total, (int64_t)blocks * 256);
    }
    if (gb) {
        ProfScope ps("k_colsum_reduce", st);
        launch_pdl(k_colsum_reduce, dim3((C + 31) / 32), dim3(256), 0, st, (const float*)ws, gb, C, (int)blocks);
    }
    return check_launch("k_lrelu_bwd");
}

// gp [B][Hi][Wi][C] (gradient of the pooled tensor), h [B][2Hi][2Wi][C] (the LeakyReLU output that was pooled)
// -> gz [B][2Hi][2Wi][C], gb[C] (optional, overwritten; needs ws of mg_colsum_workspace_bytes(C))
int mg_unpool2_lrelu_bwd_bf16(const void* gp, const void* h, void* gz, float* gb, void* ws, size_t ws_bytes,
                              int B, int Hi, int Wi, int C, mgStream stream) {
    if (!gp || !h || !gz || B <= 0 || Hi <= 0 || Wi <= 0 || C < 8 || (C & 7)) return MG_ERR_BAD_ARG;
    if (gb && (!ws || ws_bytes < mg_colsum_workspace_bytes(C))) return MG_ERR_WORKSPACE;
    const int C8 = C / 8;
    const int64_t total = (int64_t)B * Hi * Wi * C8;
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned blocks = colsum_blocks(total, C8);
    {
        ProfScope ps("k_unpool2_lrelu_bwd", st);
        launch_pdl(k_unpool2_lrelu_bwd, dim3(blocks), dim3(256), 0, st, (const __nv_bfloat16*)gp, (const __nv_bfloat16*)h, (__nv_bfloat16*)gz,
                   gb ? (float*)ws : (float*)nullptr, Hi, Wi, C8, total, (int64_t)blocks * 256);
    }
    if (gb) {
        ProfScope ps("k_colsum_reduce", st);
        launch_pdl(k_colsum_reduce, dim3((C + 31) / 32), dim3(256), 0, st, (const float*)ws, gb, C, (int)blocks);
    }
    return check_launch("k_unpool2_lrelu_bwd");
}

int mg_pixelnorm_lrelu_bwd_bf16(const void* go, const void* o, const float* inv_norm, void* gz, float* gb, void* ws, size_t ws_bytes,
                                int64_t n_pixels, int C, mgStream stream) {
    if (!go || !o || !inv_norm || !gz || n_pixels <= 0 || C < 8 || (C & 7) || C > 1024) return MG_ERR_BAD_ARG;
    if (gb && (!ws || ws_bytes < mg_colsum_workspace_bytes(C))) return MG_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    {
        ProfScope ps("k_pixelnorm_lrelu_bwd", st);
        launch_pdl(k_pixelnorm_lrelu_bwd, dim3(grid_for(n_pixels, 256, 148 * 8)), dim3(256), 0, st,
                   (const __nv_bfloat16*)go, (const __nv_bfloat16*)o, inv_norm, (__nv_bfloat16*)gz, C, n_pixels);
    }
    if (gb) {
        const int C8 = C / 8;
        const int64_t total = n_pixels * C8;
        const unsigned blocks = colsum_blocks(total, C8);
        ProfScope ps("k_colsum", st);
        launch_pdl(k_colsum, dim3(blocks), dim3(256), 0, st, (const __nv_bfloat16*)gz, (float*)ws, C8, total, (int64_t)blocks * 256);
        launch_pdl(k_colsum_reduce, dim3((C + 31) / 32), dim3(256), 0, st, (const float*)ws, gb, C, (int)blocks);
    }
    return check_launch("k_pixelnorm_lrelu_bwd");
}

int mg_pool2_bf16(const void* in, void* out, int B, int Ho, int Wo, int C, int adjoint, mgStream stream) {
    if (!in || !out || B <= 0 || Ho <= 0 || Wo <= 0 || C < 8 || (C & 7)) return MG_ERR_BAD_ARG;
    const int64_t total = (int64_t)B * Ho * Wo * (C / 8);
    cudaStream_t st = (cudaStream_t)stream;
    ProfScope ps(adjoint == 1 ? "k_unpool2" : "k_pool2", st);
    if (adjoint != 1) launch_pdl(k_pool2, dim3(grid_for(total)), dim3(256), 0, st, (const __nv_bfloat16*)in, (__nv_bfloat16*)out, Ho, Wo, C / 8, total, adjoint == 2 ? 1.0f : 0.25f);
    else launch_pdl(k_unpool2, dim3(grid_for(total)), dim3(256), 0, st, (const __nv_bfloat16*)in, (__nv_bfloat16*)out, Ho, Wo, C / 8, total);
    return check_launch("k_pool2");
}

}  // extern "C"
