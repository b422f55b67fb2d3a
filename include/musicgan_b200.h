/* musicgan_b200 -- C ABI of the B200 (sm_100a) hot path of Ipsedo/MusicGAN.
 *
 * The reference has no FFI of its own (it is pure Python on torch): the boundary a maintainer
 * binds is this header, called from `music_gan/audio/functions.py` and `music_gan/networks/*.py`
 * through ctypes (INTEGRATION.md shows the stub).  Every entry point names the reference
 * function (file:line under /root/reference/music_gan/) whose arithmetic it replaces.
 *
 * Conventions
 *   - all pointers are DEVICE pointers unless the name ends in `_host`;
 *   - every call only enqueues work on `stream` (a cudaStream_t) and never synchronises;
 *   - the caller owns every buffer; scratch is passed in (`ws`) and sized by *_workspace_bytes;
 *   - return value: 0 = MG_OK, negative = mgError (see mg_error_string); no exceptions, no
 *     hidden allocation, no CPU fallback.
 */
#ifndef MUSICGAN_B200_H
#define MUSICGAN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* mgStream; /* cudaStream_t */

enum mgError {
    MG_OK = 0,
    MG_ERR_BAD_ARG = -1,       /* null pointer, non-positive size, misaligned buffer        */
    MG_ERR_UNSUPPORTED = -2,   /* shape / parameter outside what the sm_100a kernels cover  */
    MG_ERR_WORKSPACE = -3,     /* workspace smaller than *_workspace_bytes                  */
    MG_ERR_LAUNCH = -4,        /* CUDA launch / runtime error (cudaGetLastError text logged) */
    MG_ERR_NO_DEVICE = -5      /* no sm_100 device visible                                  */
};

int mg_version(void);
const char* mg_error_string(int err);
/* last CUDA error text recorded by a failing call of this thread ("" if none) */
const char* mg_last_cuda_error(void);
int mg_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* Optional per-kernel timing (CUDA events recorded on the caller's stream around every kernel the
 * library launches).  Off by default; bench.py switches it on for the timed region to report the
 * dominant kernel's duration.  mg_profile_collect synchronises the recorded events, returns the
 * number of distinct kernels (<= max_entries) and fills name / total milliseconds / launch count,
 * then clears the record. */
void mg_profile_enable(int on);
int mg_profile_collect(int max_entries, const char** names, float* total_ms, int* launches);

/* ------------------------------------------------------------------------------------------
 * Audio transform constants: audio/constant.py:1-4
 * ---------------------------------------------------------------------------------------- */
#define MG_N_FFT 1024
#define MG_N_VEC 512
#define MG_STFT_STRIDE 256
#define MG_SAMPLE_RATE 44100

/* Integer frame / chunk plan of audio/functions.py:53-62 (T = 1 + N / hop) and :76-92
 * (T-1 difference columns, leading (T-1) % nb_vec dropped, rest split in nb_vec-wide chunks).
 * Host-only arithmetic; bit exact by construction. */
int mg_chunk_plan(int64_t n_samples, int hop, int nb_vec, int64_t* n_frames, int64_t* head, int64_t* n_chunks);

/* Host helpers that fill the two constant vectors the kernels take as arguments, computed in
 * double precision: periodic Hann window (functions.py:51) and the normalised Bark gain
 * 6*asinh(linspace(20,22050,F)/600)/||.||2 (functions.py:29-33).  The Python binding passes
 * torch's own vectors instead so that they are bit-identical to the reference's. */
void mg_fill_hann_host(float* window_host, int n_fft);
void mg_fill_bark_gain_host(float* gain_host, int n_bins);

/* ------------------------------------------------------------------------------------------
 * Forward transform  (wav -> STFT -> magnitude / instantaneous frequency chunks)
 *   replaces audio/functions.py:38-62 wav_to_stft (minus the file read) and
 *            audio/functions.py:65-94 stft_to_phase_magn (incl. :13-35 diff/unwrap/bark)
 *
 * wav        [batch][channels][n_samples] fp32, clip stride `clip_stride` elements (mono mean over
 *            channels as functions.py:49); n_fft = 1024, hop = 256, nb_vec = 512 only.
 * window     [1024] fp32 analysis window, bark_gain [512] fp32.
 * magn/ifreq [batch][n_chunks][512][512] fp32, time fastest, normalised to [-1, 1] per clip.
 * minmax     [batch][4] fp32: raw magn min, magn max, IF min, IF max of each clip (functions.py:79-82).
 * ---------------------------------------------------------------------------------------- */
size_t mg_stft_magif_workspace_bytes(int64_t n_samples, int batch);
int mg_stft_magif_f32(const float* wav, int64_t n_samples, int channels, int64_t clip_stride, int batch,
                      const float* window, const float* bark_gain,
                      float* magn, float* ifreq, float* minmax,
                      void* ws, size_t ws_bytes, mgStream stream);

/* STFT only: functions.py:38-62.  out [batch][T][512] complex64 (re,im interleaved), FRAME major
 * (the reference's (512, T) tensor is the transposed view of this memory, stride (1, 512)). */
int mg_stft_c64(const float* wav, int64_t n_samples, int channels, int64_t clip_stride, int batch,
                const float* window, float* out_c64, mgStream stream);

/* Post-FFT stage on a caller supplied complex STFT: functions.py:65-94.
 * stft: complex64 element (f, t) of clip b at stft[2*(b*batch_stride + f*stride_f + t*stride_t)].
 * Used for the staged parity protocol and by stft_to_phase_magn(). */
size_t mg_phase_magn_workspace_bytes(int64_t n_frames, int batch);
int mg_phase_magn_from_stft(const float* stft_c64, int64_t n_frames, int64_t stride_f, int64_t stride_t,
                            int64_t batch_stride, int batch, const float* bark_gain,
                            float* magn, float* ifreq, float* minmax,
                            void* ws, size_t ws_bytes, mgStream stream);

/* ------------------------------------------------------------------------------------------
 * Inverse transform (magnitude / IF images -> cumulative phase -> complex spectrum -> iSTFT)
 *   replaces audio/functions.py:97-137 magn_phase_to_wav (minus the file write :139)
 *
 * magn_phase [n_clips][imgs_per_clip][2][512][W] fp32; the images of a clip are concatenated
 *            along time (functions.py:108-109), each clip is de-normalised on its own
 *            (generate.py:61-65 passes one image per call).
 * wav        [n_clips][256 * (imgs_per_clip*W - 1)] fp32.
 * ---------------------------------------------------------------------------------------- */
size_t mg_istft_workspace_bytes(int n_clips, int imgs_per_clip, int width);
int mg_istft_from_magif_f32(const float* magn_phase, int n_clips, int imgs_per_clip, int width,
                            const float* window, const float* bark_gain,
                            float* wav, void* ws, size_t ws_bytes, mgStream stream);

/* ------------------------------------------------------------------------------------------
 * ProGAN 3x3 convolutions (stride 1, pad 1) as tcgen05 implicit GEMMs, NHWC bf16 activations, fp32
 * accumulation, fp32 master weights [Cout][Cin][3][3] packed to bf16 inside the call.
 *   replaces nn.Conv2d(3x3) of networks/generator.py:16-22,31-37 and networks/discriminator.py:15-21,26-32
 *   (+ LeakyReLU generator.py:23,38 / discriminator.py:22,33, PixelNorm layers.py:11-17 and the nearest
 *   x2 Upsample generator.py:26-29 when the corresponding flag is set).
 *
 * x [B][H(/2)][W(/2)][Cin] bf16, y [B][H][W][Cout] bf16; H, W are the OUTPUT dims.
 * flags: 1 = LeakyReLU(0.2) epilogue, 2 = PixelNorm epilogue (writes inv_norm [B][H][W] fp32 if non-null),
 *        4 = input is read through a nearest x2 upsampling, 8 = data-gradient mode: `w_f32` is still the
 *        forward weight [Cfwd_out = Cin][Cfwd_in = Cout][3][3] and y = dL/dx of the forward conv for x = dL/dy,
 *        16 = weights enter as hi + lo bf16 pairs (w = bf16(w) + bf16(w - bf16(w)), two MMA sweeps per tile): used by
 *        the forward convolutions, whose pre-activation SIGNS decide the LeakyReLU masks of the whole backward pass.
 * ---------------------------------------------------------------------------------------- */
size_t mg_conv3x3_workspace_bytes(int Cin, int Cout);
/* Pack fp32 weights once (e.g. per optimiser step) into `packed` (>= mg_conv3x3_workspace_bytes); mg_conv3x3_bf16
 * called with w_f32 == NULL then reads its `ws` argument as such a pre-packed buffer (same Cin, Cout and mode).
 * mode: bit 0 data-gradient orientation (flag 8), bit 1 hi + lo pairs (flag 16), bit 2 the layer runs with flag 2. */
int mg_conv3x3_pack_weights(const float* w_f32, int Cin, int Cout, int mode, void* packed, size_t packed_bytes, mgStream stream);
int mg_conv3x3_bf16(const void* x, const float* w_f32, const float* bias, void* y, float* inv_norm,
                    int B, int H, int W, int Cin, int Cout, int flags, void* ws, size_t ws_bytes, mgStream stream);

/* The same convolution on fp32 NHWC activations with split-bf16 operands (x = hi + lo, w = hi + lo, three MMAs per
 * K step: ~16 significand bits): the precise path of the layers whose output height is <= 32, where bf16 operand
 * rounding flips LeakyReLU masks and moves the WGAN-GP gradients by several percent (conv_split.cu).
 * x [B][H(/2)][W(/2)][Cin] fp32, y [B][H][W][Cout] fp32, y_bf16 optional bf16 copy of y (operand of
 * mg_conv3x3_wgrad_bf16), `packed` from mg_conv3x3_split_pack_weights (>= mg_conv3x3_split_workspace_bytes; mode bit 0
 * packs the data-gradient orientation, to be used with flag 8; mode bit 1 packs w = hi + mid + lo, the fp32 weight
 * exactly, to be used with flag 32: five MMAs per K step -- the critic's forward convolutions, where one flipped
 * LeakyReLU mask in a 2 x 2-pixel layer moves the penalty gradient by percents).  flags 1, 2, 4, 8 as mg_conv3x3_bf16. */
size_t mg_conv3x3_split_workspace_bytes(int Cin, int Cout);
int mg_conv3x3_split_pack_weights(const float* w_f32, int Cin, int Cout, int mode, void* packed, size_t packed_bytes, mgStream stream);
int mg_conv3x3_split_f32(const float* x, const void* packed, const float* bias, float* y, void* y_bf16, float* inv_norm,
                         int B, int H, int W, int Cin, int Cout, int flags, mgStream stream);

/* All packed copies of all weights of a network in one launch (a training step re-packs every 3x3 weight after each
 * optimiser step).  Fill one record per packed copy on the host with mg_pack_job_fill (kind 0: the layout of
 * mg_conv3x3_pack_weights(mode); kind 1: the layout of mg_conv3x3_split_pack_weights(dgrad = mode & 1); Cin / Cout as
 * in those calls), upload the array once, then call mg_pack_weights_multi after every parameter update. */
typedef struct mgPackJob {
    const float* w;      /* fp32 master weight [Cout][Cin][3][3] (device) */
    void* out;           /* packed buffer (device) */
    int cout_fwd, cin_fwd, flip, nt, parts, kind, total, reserved;
} mgPackJob;
int mg_pack_job_fill(mgPackJob* job_host, const float* w_f32, void* packed, int Cin, int Cout, int kind, int mode);
int mg_pack_weights_multi(const mgPackJob* jobs_dev, int n_jobs, int max_total, mgStream stream);

/* Weight gradient of the same convolution: dw[co][ci][ky][kx] = sum_{b,y,x} dy[b][y][x][co] *
 * xin[b][y+ky-1][x+kx-1][ci]   (xin = x, or x read through a nearest x2 upsampling when flags bit 0 is set).
 * dy [B][H][W][Cout] bf16, x [B][H(/2)][W(/2)][Cin] bf16, dw fp32 [Cout][Cin][3][3] OVERWRITTEN -- or, with flags
 * bit 1, ADDED TO (a further contribution to the same parameter's gradient) -- by a deterministic two-stage reduction
 * through `ws` (>= mg_conv3x3_wgrad_workspace_bytes). */
size_t mg_conv3x3_wgrad_workspace_bytes(int B, int H, int W, int Cin, int Cout);
int mg_conv3x3_wgrad_bf16(const void* dy, const void* x, float* dw, void* ws, size_t ws_bytes,
                          int B, int H, int W, int Cin, int Cout, int flags, mgStream stream);
/* The same launch also producing the BIAS gradient db[co] = sum_{b,y,x} dy[b][y][x][co] (fp32 [Cout]) as one more row of
 * the GEMM (an operand row of ones): no separate column-sum pass over dy.  flags bit 2: add to db instead of overwriting. */
int mg_conv3x3_wgrad_bias_bf16(const void* dy, const void* x, float* dw, float* db, void* ws, size_t ws_bytes,
                               int B, int H, int W, int Cin, int Cout, int flags, mgStream stream);

/* ------------------------------------------------------------------------------------------
 * Memory-bound layers around the convolutions (pointwise.cu).  "mask" = LeakyReLU(0.2) derivative taken
 * from the sign of a saved activation tensor (1 where > 0, else 0.2).
 *
 * mg_rgb_expand_bf16   2 -> C 1x1 conv: x [B][2][HW] fp32 planes, w [C][2], b [C]|NULL -> y [B][HW][C] bf16 NHWC;
 *                      mode 0 plain, 1 LeakyReLU(0.2), 2 multiply by mask(mask_src)
 *                      (MagPhaseLayer, networks/discriminator.py:37-50)
 * mg_rgb_project_bf16  C -> 2 1x1 conv: a [B][HW][C] bf16 (optionally masked) -> out [B][2][HW] fp32 planes;
 *                      weight row k, column c at w2[k*row_stride + c*col_stride]; bias [2]|NULL; act 1 = tanh
 *                      (ToMagnPhaseLayer, networks/generator.py:43-52, and rgb_expand's data gradient)
 * mg_rgb_wgrad_bf16    gw[c][k] += sum_p (g*mask)[p][c] x[k][p], gb[c] += sum_p (g*mask)[p][c]  (fp32, caller zeroes)
 * mg_pool2_bf16        AvgPool2d(2,2) on bf16 NHWC (networks/discriminator.py:24,131); adjoint != 0 runs its
 *                      transpose (each input pixel * 0.25 replicated to a 2x2 block); adjoint == 2: 2x2 SUM pooling (backward of
 *                      the nearest x2 upsampling, generator.py:26-29); Ho, Wo = the SMALLER dims
 * ---------------------------------------------------------------------------------------- */
int mg_rgb_expand_bf16(const float* x, const float* w, const float* b, const void* mask_src, void* y,
                       int B, int64_t HW, int C, int mode, mgStream stream);
int mg_rgb_project_bf16(const void* a, const float* w2, int row_stride, int col_stride, const float* bias, const void* mask_src,
                        float* out, int B, int64_t HW, int C, int act, mgStream stream);
int mg_rgb_wgrad_bf16(const void* g, const void* mask_src, const float* x, float* gw, float* gb, int B, int64_t HW, int C, mgStream stream);
int mg_pool2_bf16(const void* in, void* out, int B, int Ho, int Wo, int C, int adjoint, mgStream stream);
/* AvgPool2d(2,2) on fp32 NCHW planes (network input on the fade-in path, discriminator.py:130-133):
 * adjoint 0: in [n_planes][2Ho][2Wo] -> out [n_planes][Ho][Wo]; adjoint 1: its backward, in [n_planes][Ho][Wo] ->
 * out [n_planes][2Ho][2Wo] = 0.25 * in replicated.  Bit-identical to torch.nn.functional.avg_pool2d / its backward. */
int mg_pool2_planes_f32(const float* in, float* out, int64_t n_planes, int Ho, int Wo, int adjoint, mgStream stream);
/* Scratch of the per-channel sums below: one row of C partial sums per thread block, summed by a second small kernel in
 * a fixed order (deterministic, no atomics). */
size_t mg_colsum_workspace_bytes(int C);
/* LeakyReLU(0.2) backward fused with the bias gradient: gz = gy * mask(y) (bf16 NHWC), gb[c] = sum over pixels of gz
 * (fp32, OVERWRITTEN, may be NULL; needs ws of mg_colsum_workspace_bytes(C)).
 * (backward of the Conv2d + LeakyReLU pairs, discriminator.py:15-22,26-33) */
int mg_lrelu_bwd_bf16(const void* gy, const void* y, void* gz, float* gb, void* ws, size_t ws_bytes,
                      int64_t n_pixels, int C, mgStream stream);
/* Backward of LeakyReLU -> AvgPool2d(2,2) in one pass (discriminator.py:15-24, ConvBlock): gp [B][Hi][Wi][C] is the
 * gradient of the pooled tensor, h [B][2Hi][2Wi][C] the LeakyReLU output that was pooled;
 * gz = 0.25 * upsample(gp) * mask(h) (bf16 NHWC, full resolution), gb[c] = sum over pixels of gz (OVERWRITTEN, may be
 * NULL; needs ws of mg_colsum_workspace_bytes(C)).  Same values as mg_pool2_bf16(adjoint) followed by mg_lrelu_bwd_bf16. */
int mg_unpool2_lrelu_bwd_bf16(const void* gp, const void* h, void* gz, float* gb, void* ws, size_t ws_bytes,
                              int B, int Hi, int Wi, int C, mgStream stream);
/* PixelNorm + LeakyReLU backward of a generator half-block (layers.py:11-17 after generator.py:23,38):
 * go, o [n_pixels][C] bf16 (o = the normalised forward output), inv_norm [n_pixels] fp32 (from mg_conv3x3_bf16 flag 2)
 * -> gz [n_pixels][C] bf16 = gradient w.r.t. the convolution output (+bias), gb[c] = sum over pixels (OVERWRITTEN, may be
 * NULL; needs ws of mg_colsum_workspace_bytes(C)). */
int mg_pixelnorm_lrelu_bwd_bf16(const void* go, const void* o, const float* inv_norm, void* gz, float* gb, void* ws, size_t ws_bytes,
                                int64_t n_pixels, int C, mgStream stream);

/* The same memory-bound layers on fp32 NHWC activations (precise path of the low-resolution layers; same arguments,
 * every `void*` activation / mask tensor is fp32 instead of bf16). */
int mg_rgb_expand_f32(const float* x, const float* w, const float* b, const void* mask_src, void* y,
                      int B, int64_t HW, int C, int mode, mgStream stream);
int mg_rgb_project_f32(const void* a, const float* w2, int row_stride, int col_stride, const float* bias, const void* mask_src,
                       float* out, int B, int64_t HW, int C, int act, mgStream stream);
int mg_rgb_wgrad_f32(const void* g, const void* mask_src, const float* x, float* gw, float* gb, int B, int64_t HW, int C, mgStream stream);
int mg_pool2_f32(const void* in, void* out, int B, int Ho, int Wo, int C, int adjoint, mgStream stream);
int mg_lrelu_bwd_f32(const void* gy, const void* y, void* gz, float* gb, void* ws, size_t ws_bytes,
                     int64_t n_pixels, int C, mgStream stream);
int mg_unpool2_lrelu_bwd_f32(const void* gp, const void* h, void* gz, float* gb, void* ws, size_t ws_bytes,
                             int B, int Hi, int Wi, int C, mgStream stream);
int mg_pixelnorm_lrelu_bwd_f32(const void* go, const void* o, const float* inv_norm, void* gz, float* gb, void* ws, size_t ws_bytes,
                               int64_t n_pixels, int C, mgStream stream);

#ifdef __cplusplus
}
#endif
#endif /* MUSICGAN_B200_H */
