// Shared host/device helpers of the musicgan_b200 C-ABI library (sm_100a only).
#pragma once
#include <cstdlib>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/musicgan_b200.h"

namespace mg {

// fp32 roundings of numpy's pi / 2*pi: what the reference's `np.pi` scalars become when they
// meet a float32 tensor (audio/functions.py:19-22,115,120).
constexpr float kPiF = 3.14159274101257324f;       // 0x40490FDB
constexpr float kTwoPiF = 6.28318548202514648f;    // 0x40C90FDB

void set_last_cuda_error(const char* where, cudaError_t e);

inline int check_launch(const char* where) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_last_cuda_error(where, e); return MG_ERR_LAUNCH; }
    return MG_OK;
}

// RAII stage timer (no-op unless mg_profile_enable(1)): records an event pair on `st`.
struct ProfScope {
    int slot;
    cudaStream_t st;
    ProfScope(const char* name, cudaStream_t st);
    ~ProfScope();
};

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// SM count of the CURRENT device (cached per device: the Python API accepts any device)
inline int current_sm_count() {
    static int cache[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (!cache[dev]) {
        int n = 0;
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        cache[dev] = n > 0 ? n : 148;
    }
    return cache[dev];
}

// ---- programmatic dependent launch ------------------------------------------------------------------------------
// A training step is ~1000 short dependent kernels.  Launched with programmaticStreamSerialization the next kernel's
// CTAs may become resident and run their on-chip prologue (barrier init, TMEM allocation, index tables) while the
// previous kernel is still draining; `pdl_wait()` then blocks until that kernel has completed and its writes are
// visible.  Rules kept by every kernel launched through launch_pdl(): pdl_trigger() first, pdl_wait() executed by EVERY
// thread before its first global-memory access (reads AND writes: an output buffer may be recycled memory the previous
// kernel still reads).  Without the launch attribute both instructions are no-ops.  MG_PDL=0 switches the attribute off.
#if defined(__CUDACC__)
// threadIdx.x read exactly once: under register pressure the compiler otherwise re-reads the special register (S2R, ~20
// cycles) inside the per-tile loops of the warp-specialised kernels instead of keeping it live
__device__ __forceinline__ int tid_x_once() { int t; asm volatile("mov.u32 %0, %%tid.x;" : "=r"(t)); return t; }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
    static const bool on = !(getenv("MG_PDL") && atoi(getenv("MG_PDL")) == 0);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = on ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}
#endif

// Monotone int key of a float (for atomicMin / atomicMax on floats of either sign).
__device__ __forceinline__ int float_key(float x) {
    int k = __float_as_int(x);
    return k >= 0 ? k : k ^ 0x7fffffff;
}
__device__ __forceinline__ float key_float(int k) {
    return __int_as_float(k >= 0 ? k : k ^ 0x7fffffff);
}

__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// torch.remainder(x, m) for m > 0 on float32: fmodf + sign fix (ATen remainder kernel).
__device__ __forceinline__ float remainder_pos(float x, float m) {
    float r = fmodf(x, m);
    if (r != 0.0f && r < 0.0f) r = __fadd_rn(r, m);
    return r;
}

// One element of the reference's unwrap() adjustment, audio/functions.py:17-22, in the exact
// float32 operation order (no FMA contraction: every op is an explicit _rn intrinsic):
//   dphi   = phi[t] - phi[t-1]
//   dphi_m = ((dphi + pi) % 2pi) - pi ;  dphi_m = pi where (dphi_m == -pi) & (dphi > 0)
//   adj    = dphi_m - dphi ;             adj = 0 where |dphi| < pi
__device__ __forceinline__ float unwrap_adjust(float phi_prev, float phi) {
    const float d = __fsub_rn(phi, phi_prev);
    if (fabsf(d) < kPiF) return 0.0f;
    const float x = __fadd_rn(d, kPiF);
    float r;
    if (x >= 0.0f && x < 2.0f * kTwoPiF) {
        // fmodf shortcut, exact: for m <= x < 2m, x - m is exactly representable (Sterbenz)
        r = x >= kTwoPiF ? __fsub_rn(x, kTwoPiF) : x;
    } else {
        r = remainder_pos(x, kTwoPiF);
    }
    float dm = __fsub_rn(r, kPiF);
    if (dm == -kPiF && d > 0.0f) dm = kPiF;
    return __fsub_rn(dm, d);
}

}  // namespace mg
