// Library-level entry points of the musicgan_b200 C ABI (version, errors, host table helpers).
#include "common.cuh"

#include <cmath>
#include <cstring>
#include <vector>

namespace mg {
static thread_local char g_last_error[256] = "";
void set_last_cuda_error(const char* where, cudaError_t e) {
    snprintf(g_last_error, sizeof(g_last_error), "%s: %s", where, cudaGetErrorString(e));
}

// ---- stage profiler ----------------------------------------------------------------------------
static bool g_prof_on = false;
struct ProfRec { const char* name; cudaEvent_t a, b; };
static std::vector<ProfRec> g_prof;
static std::vector<cudaEvent_t> g_event_pool;
static cudaEvent_t take_event() {
    if (!g_event_pool.empty()) { cudaEvent_t e = g_event_pool.back(); g_event_pool.pop_back(); return e; }
    cudaEvent_t e; cudaEventCreate(&e); return e;
}
ProfScope::ProfScope(const char* name, cudaStream_t s) : slot(-1), st(s) {
    if (!g_prof_on) return;
    ProfRec r{name, take_event(), take_event()};
    cudaEventRecord(r.a, st);
    slot = (int)g_prof.size();
    g_prof.push_back(r);
}
ProfScope::~ProfScope() {
    if (slot >= 0) cudaEventRecord(g_prof[slot].b, st);
}
}  // namespace mg

extern "C" {

void mg_profile_enable(int on) { mg::g_prof_on = on != 0; }

int mg_profile_collect(int max_entries, const char** names, float* total_ms, int* launches) {
    int n = 0;
    for (auto& r : mg::g_prof) {
        cudaEventSynchronize(r.b);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, r.a, r.b);
        int k = 0;
        for (; k < n; ++k) if (strcmp(names[k], r.name) == 0) break;
        if (k == n) { if (n >= max_entries) continue; names[n] = r.name; total_ms[n] = 0.f; launches[n] = 0; ++n; }
        total_ms[k] += ms; launches[k] += 1;
        mg::g_event_pool.push_back(r.a); mg::g_event_pool.push_back(r.b);
    }
    mg::g_prof.clear();
    return n;
}

int mg_version(void) { return 100; }   // 0.1.0

const char* mg_error_string(int err) {
    switch (err) {
        case MG_OK: return "ok";
        case MG_ERR_BAD_ARG: return "bad argument (null pointer, non-positive size or misaligned buffer)";
        case MG_ERR_UNSUPPORTED: return "unsupported shape or parameter for the sm_100a kernels";
        case MG_ERR_WORKSPACE: return "workspace too small";
        case MG_ERR_LAUNCH: return "CUDA launch / runtime error (see mg_last_cuda_error)";
        case MG_ERR_NO_DEVICE: return "no sm_100 CUDA device";
        default: return "unknown error";
    }
}

const char* mg_last_cuda_error(void) { return mg::g_last_error; }

int mg_device_info(int* sm_count, int* cc_major, int* cc_minor) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) { mg::set_last_cuda_error("cudaGetDevice", e); return MG_ERR_NO_DEVICE; }
    cudaDeviceProp p;
    e = cudaGetDeviceProperties(&p, dev);
    if (e != cudaSuccess) { mg::set_last_cuda_error("cudaGetDeviceProperties", e); return MG_ERR_NO_DEVICE; }
    if (sm_count) *sm_count = p.multiProcessorCount;
    if (cc_major) *cc_major = p.major;
    if (cc_minor) *cc_minor = p.minor;
    return MG_OK;
}

// periodic Hann: 0.5 - 0.5 cos(2 pi n / N)   (reference audio/functions.py:51, th.hann_window)
void mg_fill_hann_host(float* window_host, int n_fft) {
    for (int n = 0; n < n_fft; ++n)
        window_host[n] = (float)(0.5 - 0.5 * std::cos(2.0 * M_PI * (double)n / (double)n_fft));
}

// reference audio/functions.py:29-33
void mg_fill_bark_gain_host(float* gain_host, int n_bins) {
    std::vector<double> s(n_bins);
    double nrm = 0.0;
    for (int i = 0; i < n_bins; ++i) {
        const double hz = n_bins > 1 ? 20.0 + (22050.0 - 20.0) * (double)i / (double)(n_bins - 1) : 20.0;
        s[i] = 6.0 * std::asinh(hz / 600.0);
        nrm += s[i] * s[i];
    }
    nrm = std::sqrt(nrm);
    for (int i = 0; i < n_bins; ++i) gain_host[i] = (float)(s[i] / nrm);
}

}  // extern "C"
