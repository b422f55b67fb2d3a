// Device-resident constant tables shared by the transform kernels (defined in audio_forward.cu).
#pragma once
#include "common.cuh"
#include "fft512.cuh"

namespace mg {
struct DeviceTables {
    FftTables fft;
    float2 w1024[512];   // exp(-2 pi i k / 1024)
};
cudaError_t ensure_tables();                 // uploads once per process (one process per GPU)
const DeviceTables* device_tables_ptr();
void launch_init_keys(int* keys, int n_clips, cudaStream_t st);   // min slots = INT_MAX, max slots = INT_MIN
}  // namespace mg
