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

// Activation element type of a kernel: bf16 (the default storage) or fp32 (the precise path of the low-resolution
// layers, conv_split.cu).  `base` points at the tensor, `i` counts 8-element chunks.
template <typename T> struct Act8;
template <> struct Act8<__nv_bfloat16> {
    template <bool kStream = false>
    static __device__ __forceinline__ void ld(const __nv_bfloat16* base, int64_t i, float (&v)[8]) {
        Pack8 p;
        p.u = kStream ? __ldcs(reinterpret_cast<const uint4*>(base) + i) : __ldg(reinterpret_cast<const uint4*>(base) + i);
#pragma unroll
        for (int j = 0; j < 4; ++j) { const float2 f = __bfloat1622float2(p.h[j]); v[2 * j] = f.x; v[2 * j + 1] = f.y; }
    }
    static __device__ __forceinline__ void st(__nv_bfloat16* base, int64_t i, const float (&v)[8]) {
        Pack8 p;
#pragma unroll
        for (int j = 0; j < 4; ++j) p.h[j] = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
        reinterpret_cast<uint4*>(base)[i] = p.u;
    }
};
template <> struct Act8<float> {
    template <bool kStream = false>
    static __device__ __forceinline__ void ld(const float* base, int64_t i, float (&v)[8]) {
        const float4* s = reinterpret_cast<const float4*>(base) + 2 * i;
        const float4 a = kStream ? __ldcs(s) : __ldg(s), b = kStream ? __ldcs(s + 1) : __ldg(s + 1);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    }
    static __device__ __forceinline__ void st(float* base, int64_t i, const float (&v)[8]) {
        float4* d = reinterpret_cast<float4*>(base) + 2 * i;
        d[0] = make_float4(v[0], v[1], v[2], v[3]);
        d[1] = make_float4(v[4], v[5], v[6], v[7]);
    }
};

// ---- 2 -> C ------------------------------------------------------------------------------------
// x [B][2][HW] fp32, w [C][2], b [C] or null, mask_src [B][HW][C] bf16 or null, y [B][HW][C] bf16
// mode: 0 none, 1 LeakyReLU(0.2) on the result, 2 multiply the result by mask(mask_src)
template <typename T>
__global__ void __launch_bounds__(256)
k_rgb_expand(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b,
             const T* __restrict__ mask_src, T* __restrict__ y, int64_t HW, int C, int mode, int64_t total) {
    pdl_trigger();
    pdl_wait();
    extern __shared__ float sw[];           // [C][2] then [C] bias
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) sw[i] = w[i];
    for (int i = threadIdx.x; i < C; i += blockDim.x) sw[2 * C + i] = b ? b[i] : 0.0f;
    __syncthreads();
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < total; p += (int64_t)gridDim.x * blockDim.x) {
        const int64_t bi = p / HW, q = p - bi * HW;
        const float x0 = x[(bi * 2) * HW + q], x1 = x[(bi * 2 + 1) * HW + q];
        for (int c0 = 0; c0 < C; c0 += 8) {
            float o[8], m[8];
            if (mode == 2) Act8<T>::ld(mask_src + p * C, c0 >> 3, m);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int c = c0 + j;
                float v = fmaf(sw[2 * c], x0, fmaf(sw[2 * c + 1], x1, sw[2 * C + c]));
                if (mode == 1) v = lrelu02(v);
                if (mode == 2) v *= m[j] > 0.0f ? 1.0f : 0.2f;
                o[j] = v;
            }
            Act8<T>::st(y + p * C, c0 >> 3, o);
        }
    }
}

// ---- C -> 2 ------------------------------------------------------------------------------------
// a [B][HW][C] bf16, w2 = two rows of C weights: row k at w2[k * row_stride + c * col_stride] (so both a [2][C]
// forward weight and the transpose of a [C][2] weight can be passed), bias [2] or null, mask_src or null,
// out [B][2][HW] fp32.  act: 0 none, 1 tanh
template <typename T>
__global__ void __launch_bounds__(256)
k_rgb_project(const T* __restrict__ a, const float* __restrict__ w2, int row_stride, int col_stride,
              const float* __restrict__ bias, const T* __restrict__ mask_src, float* __restrict__ out,
              int64_t HW, int C, int act, int64_t total) {
    pdl_trigger();
    pdl_wait();
    extern __shared__ float sw[];           // [2][C]
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) { const int k = i / C, c = i - k * C; sw[i] = w2[k * row_stride + c * col_stride]; }
    __syncthreads();
    const float b0 = bias ? bias[0] : 0.0f, b1 = bias ? bias[1] : 0.0f;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < total; p += (int64_t)gridDim.x * blockDim.x) {
        const int64_t bi = p / HW, q = p - bi * HW;
        float s0 = b0, s1 = b1;
        for (int c0 = 0; c0 < C; c0 += 8) {
            float v[8], m[8];
            Act8<T>::ld(a + p * C, c0 >> 3, v);
            if (mask_src) Act8<T>::ld(mask_src + p * C, c0 >> 3, m);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float f = v[j];
                if (mask_src) f *= m[j] > 0.0f ? 1.0f : 0.2f;
                s0 = fmaf(f, sw[c0 + j], s0);
                s1 = fmaf(f, sw[C + c0 + j], s1);
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
template <typename T>
__global__ void __launch_bounds__(256)
k_rgb_wgrad(const T* __restrict__ g, const T* __restrict__ mask_src, const float* __restrict__ x,
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
        float v[8], m[8];
        Act8<T>::ld(g + p * C + c0, 0, v);
        if (mask_src) Act8<T>::ld(mask_src + p * C + c0, 0, m);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float f = v[j];
            if (mask_src) f *= m[j] > 0.0f ? 1.0f : 0.2f;
            acc[j][0] = fmaf(f, x0, acc[j][0]); acc[j][1] = fmaf(f, x1, acc[j][1]); acc[j][2] += f;
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
template <typename T>
__global__ void __launch_bounds__(256)
k_pool2(const T* __restrict__ in, T* __restrict__ out, int Ho, int Wo, int C8, int64_t total, float scale) {
    pdl_trigger();
    pdl_wait();
    // one thread = 8 channels of one output pixel
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C8);
        int64_t r = i / C8;
        const int ox = (int)(r % Wo); r /= Wo;
        const int oy = (int)(r % Ho);
        const int64_t b = r / Ho;
        const int64_t s0 = ((b * 2 * Ho + 2 * oy) * (int64_t)(2 * Wo) + 2 * ox) * C8 + c;
        float v00[8], v01[8], v10[8], v11[8], o[8];
        Act8<T>::ld(in, s0, v00); Act8<T>::ld(in, s0 + C8, v01);
        Act8<T>::ld(in, s0 + (int64_t)2 * Wo * C8, v10); Act8<T>::ld(in, s0 + (int64_t)2 * Wo * C8 + C8, v11);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = scale * ((v00[j] + v01[j]) + (v10[j] + v11[j]));
        Act8<T>::st(out, i, o);
    }
}

// adjoint: out[b][2y+i][2x+j][c] = 0.25 * in[b][y][x][c]
template <typename T>
__global__ void __launch_bounds__(256)
k_unpool2(const T* __restrict__ in, T* __restrict__ out, int Hi, int Wi, int C8, int64_t total) {
    pdl_trigger();
    pdl_wait();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C8);
        int64_t r = i / C8;
        const int x = (int)(r % Wi); r /= Wi;
        const int y = (int)(r % Hi);
        const int64_t b = r / Hi;
        float v[8];
        Act8<T>::ld(in, i, v);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] *= 0.25f;
        const int64_t d0 = ((b * 2 * Hi + 2 * y) * (int64_t)(2 * Wi) + 2 * x) * C8 + c;
        Act8<T>::st(out, d0, v); Act8<T>::st(out, d0 + C8, v);
        Act8<T>::st(out, d0 + (int64_t)2 * Wi * C8, v); Act8<T>::st(out, d0 + (int64_t)2 * Wi * C8 + C8, v);
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
template <typename T>
__global__ void __launch_bounds__(256)
k_lrelu_bwd(const T* __restrict__ gy, const T* __restrict__ y, T* __restrict__ gz,
            float* __restrict__ part, int C8, int64_t total, int64_t stride) {
    pdl_trigger();
    pdl_wait();
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.0f;
    const int64_t first = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    // 4 independent loads per operand in flight per thread (memory-level parallelism)
    int64_t i = first;
    for (; i + 3 * stride < total; i += 4 * stride) {
        float g[4][8], m[4][8];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            Act8<T>::template ld<true>(gy, i + u * stride, g[u]);
            Act8<T>::template ld<true>(y, i + u * stride, m[u]);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                g[u][j] *= m[u][j] > 0.0f ? 1.0f : 0.2f;
                acc[j] += g[u][j];
            }
            Act8<T>::st(gz, i + u * stride, g[u]);
        }
    }
    for (; i < total; i += stride) {
        float g[8], m[8];
        Act8<T>::ld(gy, i, g);
        Act8<T>::ld(y, i, m);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            g[j] *= m[j] > 0.0f ? 1.0f : 0.2f;
            acc[j] += g[j];
        }
        Act8<T>::st(gz, i, g);
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
template <typename T>
__global__ void __launch_bounds__(256)
k_unpool2_lrelu_bwd(const T* __restrict__ gp, const T* __restrict__ h, T* __restrict__ gz,
                    float* __restrict__ part, int Hi, int Wi, int C8, int64_t total, int64_t stride) {
    pdl_trigger();
    pdl_wait();
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.0f;
    const int64_t first = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t row = (int64_t)2 * Wi * C8;                 // one full-resolution image row, in 8-channel chunks
    for (int64_t i = first; i < total; i += stride) {
        const int c = (int)(i % C8);
        int64_t r = i / C8;
        const int x = (int)(r % Wi); r /= Wi;
        const int y = (int)(r % Hi);
        const int64_t b = r / Hi;
        const int64_t o0 = ((b * 2 * Hi + 2 * y) * (int64_t)(2 * Wi) + 2 * x) * C8 + c;
        float g[8], m[4][8];
        Act8<T>::ld(gp, i, g);
        Act8<T>::template ld<true>(h, o0, m[0]); Act8<T>::template ld<true>(h, o0 + C8, m[1]);
        Act8<T>::template ld<true>(h, o0 + row, m[2]); Act8<T>::template ld<true>(h, o0 + row + C8, m[3]);
#pragma unroll
        for (int j = 0; j < 8; ++j) g[j] *= 0.25f;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                m[q][j] = g[j] * (m[q][j] > 0.0f ? 1.0f : 0.2f);
                acc[j] += m[q][j];
            }
        }
        Act8<T>::st(gz, o0, m[0]); Act8<T>::st(gz, o0 + C8, m[1]);
        Act8<T>::st(gz, o0 + row, m[2]); Act8<T>::st(gz, o0 + row + C8, m[3]);
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
template <typename T>
__global__ void __launch_bounds__(256)
k_pixelnorm_lrelu_bwd(const T* __restrict__ go, const T* __restrict__ o, const float* __restrict__ inv,
                      T* __restrict__ gz, int C, int64_t n_pixels) {
    pdl_trigger();
    pdl_wait();
    const int C8 = C >> 3;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n_pixels; p += (int64_t)gridDim.x * blockDim.x) {
        float dot = 0.0f;
        for (int c = 0; c < C8; ++c) {
            float a[8], b[8];
            Act8<T>::ld(go + p * C, c, a); Act8<T>::ld(o + p * C, c, b);
#pragma unroll
            for (int j = 0; j < 8; ++j) dot = fmaf(a[j], b[j], dot);
        }
        dot /= (float)C;
        const float s = inv[p];
        for (int c = 0; c < C8; ++c) {
            float a[8], b[8];
            Act8<T>::ld(go + p * C, c, a); Act8<T>::ld(o + p * C, c, b);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float t = (a[j] - b[j] * dot) * s;
                a[j] = t * (b[j] > 0.0f ? 1.0f : 0.2f);
            }
            Act8<T>::st(gz + p * C, c, a);
        }
    }
}

// part[blockIdx.x][c] = this block's sum over its pixels of g[.,c]  (NHWC); same thread <-> channel-chunk ownership
// as k_lrelu_bwd
template <typename T>
__global__ void __launch_bounds__(256)
k_colsum(const T* __restrict__ g, float* __restrict__ part, int C8, int64_t total, int64_t stride) {
    pdl_trigger();
    pdl_wait();
    __shared__ float s_t[256 * 9];
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.0f;
    const int64_t first = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (int64_t i = first; i < total; i += stride) {
        float v[8];
        Act8<T>::ld(g, i, v);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += v[j];
    }
    block_colsum_to_row(acc, s_t, part + (size_t)blockIdx.x * C8 * 8, first, C8);
}

static unsigned grid_for(int64_t items, int per_block = 256, int cap = 148 * 16) {
    int64_t g = (items + per_block - 1) / per_block;
    return (unsigned)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace mg

using namespace mg;

// ---- host side, templated on the activation element type (bf16 | fp32) -----------------------------------------
template <typename T>
static int rgb_expand_impl(const float* x, const float* w, const float* b, const void* mask_src, void* y,
                           int B, int64_t HW, int C, int mode, mgStream stream) {
    if (!x || !w || !y || B <= 0 || HW <= 0 || C < 8 || (C & 7) || C > 512) return MG_ERR_BAD_ARG;
    if (mode == 2 && !mask_src) return MG_ERR_BAD_ARG;
    const int64_t total = (int64_t)B * HW;
    cudaStream_t st = (cudaStream_t)stream;
    ProfScope ps("k_rgb_expand", st);
    launch_pdl(k_rgb_expand<T>, dim3(grid_for(total)), dim3(256), 3 * C * sizeof(float), st, x, w, b, (const T*)mask_src, (T*)y, HW, C, mode, total);
    return check_launch("k_rgb_expand");
}

template <typename T>
static int rgb_project_impl(const void* a, const float* w2, int row_stride, int col_stride, const float* bias, const void* mask_src,
                            float* out, int B, int64_t HW, int C, int act, mgStream stream) {
    if (!a || !w2 || !out || B <= 0 || HW <= 0 || C < 8 || (C & 7) || C > 512) return MG_ERR_BAD_ARG;
    const int64_t total = (int64_t)B * HW;
    cudaStream_t st = (cudaStream_t)stream;
    ProfScope ps("k_rgb_project", st);
    launch_pdl(k_rgb_project<T>, dim3(grid_for(total)), dim3(256), 2 * C * sizeof(float), st, (const T*)a, w2, row_stride, col_stride, bias,
               (const T*)mask_src, out, HW, C, act, total);
    return check_launch("k_rgb_project");
}

template <typename T>
static int rgb_wgrad_impl(const void* g, const void* mask_src, const float* x, float* gw, float* gb, int B, int64_t HW, int C, mgStream stream) {
    if (!g || !x || !gw || B <= 0 || HW <= 0 || C < 8 || (C & 7) || C > 512) return MG_ERR_BAD_ARG;
    const int64_t total = (int64_t)B * HW;
    cudaStream_t st = (cudaStream_t)stream;
    ProfScope ps("k_rgb_wgrad", st);
    const unsigned gx = grid_for(total, 256 * 8, 148 * 2);      // every block ends with 24 same-address atomics: keep the tail short
    launch_pdl(k_rgb_wgrad<T>, dim3(gx, C / 8), dim3(256), 0, st, (const T*)g, (const T*)mask_src, x, gw, gb, HW, C, total);
    return check_launch("k_rgb_wgrad");
}

template <typename T>
static int lrelu_bwd_impl(const void* gy, const void* y, void* gz, float* gb, void* ws, size_t ws_bytes, int64_t n_pixels, int C, mgStream stream) {
    if (!gy || !y || !gz || n_pixels <= 0 || C < 8 || (C & 7)) return MG_ERR_BAD_ARG;
    if (gb && (!ws || ws_bytes < mg_colsum_workspace_bytes(C))) return MG_ERR_WORKSPACE;
    const int C8 = C / 8;
    const int64_t total = n_pixels * C8;
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned blocks = colsum_blocks(total, C8);
    {
        ProfScope ps("k_lrelu_bwd", st);
        launch_pdl(k_lrelu_bwd<T>, dim3(blocks), dim3(256), 0, st, (const T*)gy, (const T*)y, (T*)gz,
                   gb ? (float*)ws : (float*)nullptr, C8, total, (int64_t)blocks * 256);
    }
    if (gb) {
        ProfScope ps("k_colsum_reduce", st);
        launch_pdl(k_colsum_reduce, dim3((C + 31) / 32), dim3(256), 0, st, (const float*)ws, gb, C, (int)blocks);
    }
    return check_launch("k_lrelu_bwd");
}

template <typename T>
static int unpool2_lrelu_bwd_impl(const void* gp, const void* h, void* gz, float* gb, void* ws, size_t ws_bytes,
                                  int B, int Hi, int Wi, int C, mgStream stream) {
    if (!gp || !h || !gz || B <= 0 || Hi <= 0 || Wi <= 0 || C < 8 || (C & 7)) return MG_ERR_BAD_ARG;
    if (gb && (!ws || ws_bytes < mg_colsum_workspace_bytes(C))) return MG_ERR_WORKSPACE;
    const int C8 = C / 8;
    const int64_t total = (int64_t)B * Hi * Wi * C8;
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned blocks = colsum_blocks(total, C8);
    {
        ProfScope ps("k_unpool2_lrelu_bwd", st);
        launch_pdl(k_unpool2_lrelu_bwd<T>, dim3(blocks), dim3(256), 0, st, (const T*)gp, (const T*)h, (T*)gz,
                   gb ? (float*)ws : (float*)nullptr, Hi, Wi, C8, total, (int64_t)blocks * 256);
    }
    if (gb) {
        ProfScope ps("k_colsum_reduce", st);
        launch_pdl(k_colsum_reduce, dim3((C + 31) / 32), dim3(256), 0, st, (const float*)ws, gb, C, (int)blocks);
    }
    return check_launch("k_unpool2_lrelu_bwd");
}

template <typename T>
static int pixelnorm_lrelu_bwd_impl(const void* go, const void* o, const float* inv_norm, void* gz, float* gb, void* ws, size_t ws_bytes,
                                    int64_t n_pixels, int C, mgStream stream) {
    if (!go || !o || !inv_norm || !gz || n_pixels <= 0 || C < 8 || (C & 7) || C > 1024) return MG_ERR_BAD_ARG;
    if (gb && (!ws || ws_bytes < mg_colsum_workspace_bytes(C))) return MG_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    {
        ProfScope ps("k_pixelnorm_lrelu_bwd", st);
        launch_pdl(k_pixelnorm_lrelu_bwd<T>, dim3(grid_for(n_pixels, 256, 148 * 8)), dim3(256), 0, st,
                   (const T*)go, (const T*)o, inv_norm, (T*)gz, C, n_pixels);
    }
    if (gb) {
        const int C8 = C / 8;
        const int64_t total = n_pixels * C8;
        const unsigned blocks = colsum_blocks(total, C8);
        ProfScope ps("k_colsum", st);
        launch_pdl(k_colsum<T>, dim3(blocks), dim3(256), 0, st, (const T*)gz, (float*)ws, C8, total, (int64_t)blocks * 256);
        launch_pdl(k_colsum_reduce, dim3((C + 31) / 32), dim3(256), 0, st, (const float*)ws, gb, C, (int)blocks);
    }
    return check_launch("k_pixelnorm_lrelu_bwd");
}

template <typename T>
static int pool2_impl(const void* in, void* out, int B, int Ho, int Wo, int C, int adjoint, mgStream stream) {
    if (!in || !out || B <= 0 || Ho <= 0 || Wo <= 0 || C < 8 || (C & 7)) return MG_ERR_BAD_ARG;
    const int64_t total = (int64_t)B * Ho * Wo * (C / 8);
    cudaStream_t st = (cudaStream_t)stream;
    ProfScope ps(adjoint == 1 ? "k_unpool2" : "k_pool2", st);
    if (adjoint != 1) launch_pdl(k_pool2<T>, dim3(grid_for(total)), dim3(256), 0, st, (const T*)in, (T*)out, Ho, Wo, C / 8, total, adjoint == 2 ? 1.0f : 0.25f);
    else launch_pdl(k_unpool2<T>, dim3(grid_for(total)), dim3(256), 0, st, (const T*)in, (T*)out, Ho, Wo, C / 8, total);
    return check_launch("k_pool2");
}

extern "C" {

size_t mg_colsum_workspace_bytes(int C) {
    return C > 0 ? align_up((size_t)kColsumMaxBlocks * C * sizeof(float), 256) : 0;
}

int mg_rgb_expand_bf16(const float* x, const float* w, const float* b, const void* mask_src, void* y,
                       int B, int64_t HW, int C, int mode, mgStream stream) {
    return rgb_expand_impl<__nv_bfloat16>(x, w, b, mask_src, y, B, HW, C, mode, stream);
}
int mg_rgb_expand_f32(const float* x, const float* w, const float* b, const void* mask_src, void* y,
                      int B, int64_t HW, int C, int mode, mgStream stream) {
    return rgb_expand_impl<float>(x, w, b, mask_src, y, B, HW, C, mode, stream);
}
int mg_rgb_project_bf16(const void* a, const float* w2, int row_stride, int col_stride, const float* bias, const void* mask_src,
                        float* out, int B, int64_t HW, int C, int act, mgStream stream) {
    return rgb_project_impl<__nv_bfloat16>(a, w2, row_stride, col_stride, bias, mask_src, out, B, HW, C, act, stream);
}
int mg_rgb_project_f32(const void* a, const float* w2, int row_stride, int col_stride, const float* bias, const void* mask_src,
                       float* out, int B, int64_t HW, int C, int act, mgStream stream) {
    return rgb_project_impl<float>(a, w2, row_stride, col_stride, bias, mask_src, out, B, HW, C, act, stream);
}
int mg_rgb_wgrad_bf16(const void* g, const void* mask_src, const float* x, float* gw, float* gb, int B, int64_t HW, int C, mgStream stream) {
    return rgb_wgrad_impl<__nv_bfloat16>(g, mask_src, x, gw, gb, B, HW, C, stream);
}
int mg_rgb_wgrad_f32(const void* g, const void* mask_src, const float* x, float* gw, float* gb, int B, int64_t HW, int C, mgStream stream) {
    return rgb_wgrad_impl<float>(g, mask_src, x, gw, gb, B, HW, C, stream);
}

int mg_pool2_planes_f32(const float* in, float* out, int64_t n_planes, int Ho, int Wo, int adjoint, mgStream stream) {
    if (!in || !out || n_planes <= 0 || Ho <= 0 || Wo <= 0) return MG_ERR_BAD_ARG;
    const int64_t total = n_planes * Ho * Wo;
    cudaStream_t st = (cudaStream_t)stream;
    ProfScope ps("k_pool2_planes", st);
    launch_pdl(k_pool2_planes, dim3(grid_for(total)), dim3(256), 0, st, in, out, Ho, Wo, total, adjoint ? 1 : 0);
    return check_launch("k_pool2_planes");
}

// gb (optional, OVERWRITTEN) needs `ws` of mg_colsum_workspace_bytes(C)
int mg_lrelu_bwd_bf16(const void* gy, const void* y, void* gz, float* gb, void* ws, size_t ws_bytes, int64_t n_pixels, int C, mgStream stream) {
    return lrelu_bwd_impl<__nv_bfloat16>(gy, y, gz, gb, ws, ws_bytes, n_pixels, C, stream);
}
int mg_lrelu_bwd_f32(const void* gy, const void* y, void* gz, float* gb, void* ws, size_t ws_bytes, int64_t n_pixels, int C, mgStream stream) {
    return lrelu_bwd_impl<float>(gy, y, gz, gb, ws, ws_bytes, n_pixels, C, stream);
}
int mg_unpool2_lrelu_bwd_bf16(const void* gp, const void* h, void* gz, float* gb, void* ws, size_t ws_bytes,
                              int B, int Hi, int Wi, int C, mgStream stream) {
    return unpool2_lrelu_bwd_impl<__nv_bfloat16>(gp, h, gz, gb, ws, ws_bytes, B, Hi, Wi, C, stream);
}
int mg_unpool2_lrelu_bwd_f32(const void* gp, const void* h, void* gz, float* gb, void* ws, size_t ws_bytes,
                             int B, int Hi, int Wi, int C, mgStream stream) {
    return unpool2_lrelu_bwd_impl<float>(gp, h, gz, gb, ws, ws_bytes, B, Hi, Wi, C, stream);
}
int mg_pixelnorm_lrelu_bwd_bf16(const void* go, const void* o, const float* inv_norm, void* gz, float* gb, void* ws, size_t ws_bytes,
                                int64_t n_pixels, int C, mgStream stream) {
    return pixelnorm_lrelu_bwd_impl<__nv_bfloat16>(go, o, inv_norm, gz, gb, ws, ws_bytes, n_pixels, C, stream);
}
int mg_pixelnorm_lrelu_bwd_f32(const void* go, const void* o, const float* inv_norm, void* gz, float* gb, void* ws, size_t ws_bytes,
                               int64_t n_pixels, int C, mgStream stream) {
    return pixelnorm_lrelu_bwd_impl<float>(go, o, inv_norm, gz, gb, ws, ws_bytes, n_pixels, C, stream);
}
int mg_pool2_bf16(const void* in, void* out, int B, int Ho, int Wo, int C, int adjoint, mgStream stream) {
    return pool2_impl<__nv_bfloat16>(in, out, B, Ho, Wo, C, adjoint, stream);
}
int mg_pool2_f32(const void* in, void* out, int B, int Ho, int Wo, int C, int adjoint, mgStream stream) {
    return pool2_impl<float>(in, out, B, Ho, Wo, C, adjoint, stream);
}

}  // extern "C"
