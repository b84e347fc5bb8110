/* vnd_b200.h — C ABI of the B200-native velvet-noise decorrelation hot path.
 *
 * This is the drop-in boundary.  The reference (ckonst/VNDecorrelate) has no FFI; its operator
 * API is Python duck typing (src/vndecorrelate/decorrelation.py:31-59).  Every entry point below
 * names the reference function whose arithmetic it reproduces; INTEGRATION.md shows the ctypes
 * binding a maintainer of the reference would add.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes.  No exceptions or aborts cross this boundary: every
 *     function returns 0 (VND_OK) or a negative VND_E* code; vnd_status_string() names it and
 *     vnd_last_error() returns a thread-local detail string.
 *   - "_dev" entry points take DEVICE pointers, an explicit cudaStream_t (passed as void*) and a
 *     caller-owned workspace: they never allocate, never synchronise, and are re-entrant on
 *     distinct streams.
 *   - "_host" entry points take HOST pointers and a vnd_ctx that owns a device arena, page-locked
 *     staging (used when the caller's buffers are pageable) and streams; host<->device copies happen
 *     inside the call.
 *   - Signals are addressed as  element(t, c) = base[t * stride_t + c * stride_c]  (strides in
 *     ELEMENTS), so C-order (frames, channels) arrays (stride_t = C, stride_c = 1) and planar
 *     (channels, frames) arrays (stride_t = 1, stride_c = frames) use the same calls.
 *   - Nothing here has a CPU fallback: without a CUDA device every compute call fails with
 *     VND_ECUDA.
 */
#ifndef VND_B200_H
#define VND_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VND_ABI_VERSION 1

enum {
  VND_OK = 0,
  VND_EINVAL = -1,       /* bad argument (null pointer, negative size, bad enum)        */
  VND_ECUDA = -2,        /* CUDA runtime error; see vnd_last_error()                    */
  VND_EUNSUPPORTED = -3, /* valid request this build cannot serve                       */
  VND_ENOMEM = -4,       /* workspace / arena too small or allocation failed            */
  VND_EPROGRAM = -5      /* malformed tap program                                       */
};

/* element types of signal buffers */
enum { VND_F32 = 0, VND_F64 = 1 };

/* Accumulation order of the sparse FIR.
 * VND_ORDER_SEGMENTED : VelvetNoise.convolve (decorrelation.py:393-415) — per segment, subtract
 *                       the negative taps, add the positive taps, multiply by the segment gain,
 *                       add into the output; every step rounded separately (no FMA).
 * VND_ORDER_ASCENDING : convolve_velvet_noise (decorrelation.py:630-660) — ascending tap index,
 *                       y = fl32(y + fl(x * coef)); the product and the sum are float32 when the
 *                       signal and the FIR are both float32, float64 otherwise (numpy promotion).
 * VND_ORDER_ASCENDING_F64 : same, with a float64 FIR (e.g. VelvetNoise.FIR): always float64 steps.
 */
enum { VND_ORDER_SEGMENTED = 0, VND_ORDER_ASCENDING = 1, VND_ORDER_ASCENDING_F64 = 2 };

/* ---------------------------------------------------------------------------------------------
 * Tap program: the per-channel tap table in the layout the kernels read.
 *
 * One int32 array `words` plus `offsets[C + 1]` (channel c owns words[offsets[c] .. offsets[c+1])).
 * An EMPTY range marks an unfiltered channel: the input is copied through
 * (decorrelation.py:399-400).
 *
 * SEGMENTED channel block:   [ S,  (n_neg, n_pos, gain_bits) x S,  tap indices ... ]
 *     tap indices are listed segment by segment, negative list then positive list, each in the
 *     reference's list order (decorrelation.py:247-254, :539-542).  gain_bits is the IEEE-754
 *     bit pattern of float32(segment_envelope[s]).  S may be 0 (filtered channel with no taps:
 *     output is zero).
 * ASCENDING channel block:   [ K,  (index, coef_lo, coef_hi) x K ]
 *     coefficient as a float64 bit pattern split in two words; indices ascending.
 * Taps with index >= frames must be removed by the caller (the reference drops them:
 * decorrelation.py:404-410); `halo` passed to the launchers is 1 + the largest remaining index.
 * -------------------------------------------------------------------------------------------*/
typedef struct vnd_tap_program {
  const int32_t* words;   /* device (for _dev calls) or host (for _host calls) */
  const int32_t* offsets; /* C + 1 entries, same memory space as `words`       */
  int64_t n_words;
  int32_t channels;
  int32_t order;          /* VND_ORDER_*                                       */
  int32_t apply_gain;     /* 0: envelope == (1.0,), the multiply is skipped (decorrelation.py:411) */
  int32_t halo;           /* 1 + max tap index over all channels (0 if no taps) */
  int32_t max_channel_words; /* longest channel block, in words (sizes the shared-memory copy) */
} vnd_tap_program;

/* A strided signal view. */
typedef struct vnd_signal {
  void* data;
  int64_t frames;
  int32_t channels;
  int32_t dtype;     /* VND_F32 or VND_F64 */
  int64_t stride_t;  /* elements */
  int64_t stride_c;  /* elements */
} vnd_signal;

/* Post-FIR stages of VelvetNoise.decorrelate (decorrelation.py:433-440) and the Haas stage of a
 * SignalChain (decorrelation.py:202-230, LR mode), fused into the FIR pass. */
typedef struct vnd_epilogue {
  int32_t ms_encode;      /* 1: encode_signal_to_side_channel (utils/dsp.py:40-63); needs 2 channels */
  int32_t use_width;      /* 1: apply_stereo_width (utils/dsp.py:21-37)                              */
  double width;
  int32_t rms_normalize;  /* 1: rms_normalize DUAL_MONO (utils/dsp.py:87-109), numpy axis-0 order    */
  int32_t haas_delay;     /* >= 0 frames; output has frames + haas_delay frames                      */
  int32_t haas_channel;   /* channel that is delayed                                                 */
} vnd_epilogue;

/* ------------------------------------------------------------------------------ housekeeping */
int vnd_abi_version(void);
const char* vnd_version(void);
const char* vnd_status_string(int status);
const char* vnd_last_error(void);
int vnd_device_count(int* count);

/* ------------------------------------------------------------------------------ device API */

/* Sparse velvet-noise FIR, y[t,c] = sum_k coef_k * x[t + i_k, c], terms past the end dropped.
 * Replaces VelvetNoise.convolve (decorrelation.py:393-415) or convolve_velvet_noise
 * (decorrelation.py:630-660) according to taps->order.  x: F32 or F64; y: F32, same frames,
 * taps->channels channels.  With F64 input every accumulation step is done in float64 and rounded
 * to float32, which is what numpy does for `f32_array -= f64_array`. */
int vnd_sparse_fir_dev(const vnd_signal* x, const vnd_signal* y, const vnd_tap_program* taps, void* stream);

/* Bytes of workspace vnd_vn_decorrelate_dev needs for the given shape/epilogue. */
int vnd_vn_decorrelate_workspace(int64_t frames, int32_t channels, const vnd_epilogue* ep, size_t* bytes);

/* Fused VelvetNoise.decorrelate (+ optional LR Haas delay): FIR -> M/S encode -> width ->
 * RMS gain -> Haas placement.  x: F32 (already cast, decorrelation.py:426); a mono signal is
 * passed with channels = 2 and stride_c = 0 (mono_to_stereo, decorrelation.py:428-429).
 * out: F32 or F64 with frames + ep->haas_delay frames.  Without rms_normalize this is ONE kernel:
 * one read and one write per sample. */
int vnd_vn_decorrelate_dev(const vnd_signal* x, const vnd_signal* out, const vnd_tap_program* taps,
                           const vnd_epilogue* ep, void* workspace, size_t workspace_bytes, void* stream);

/* numpy axis-0 order sum of squares per channel: strict left-to-right float32 running sum of
 * float32-rounded squares (what np.mean(np.square(a), axis=0) does on a C-order array,
 * utils/dsp.py:107-109).  sums: `channels` float32 values on the device. */
int vnd_colsumsq_seq_f32_dev(const vnd_signal* a, float* sums, void* stream);

/* HaasEffect.decorrelate (decorrelation.py:192-230): x F32 (channels = 2; mono passed with
 * stride_c = 0 and mono = 1), out F64 (frames + delay, 2).  mode_ms / width as in the reference. */
int vnd_haas_dev(const vnd_signal* x, const vnd_signal* out, int32_t delay, int32_t delayed_channel,
                 int32_t mode_ms, int32_t mono, int32_t use_width, double width, void* stream);

/* In-place stereo helpers on a (frames, 2) signal, F32 or F64 (utils/dsp.py:21-63, :124-167).
 * op: 0 LR_to_MS, 1 MS_to_LR, 2 apply_stereo_width(width), 3 encode_signal_to_side_channel(x, y)
 * (x = `dry`, y = `a`), 4 rms_normalize(dry, a) DUAL_MONO.  Op 4 sums each signal in the order numpy uses for its
 * layout: frame by frame for a C-order (frames, 2) signal, pairwise per column for a planar / Fortran-ordered one; workspace:
 * 64 bytes, plus 8 * (frames / 64 + 8) when a signal is planar. */
int vnd_stereo_op_dev(const vnd_signal* a, const vnd_signal* dry, int32_t op, double width,
                      void* workspace, size_t workspace_bytes, void* stream);

/* Normalisers and polar form beyond what VelvetNoise itself uses (utils/dsp.py:71-109, :374-422).  Signals are
 * contiguous 1-D (ndim = 1) or C-order (frames, channels) arrays (ndim = 2), F32 or F64.
 *   rms_normalize: STEREO mode and 1-D signals (both statistics with axis=None: numpy's pairwise summation order is
 *       reproduced bit for bit); 2-D DUAL_MONO keeps using vnd_stereo_op op 4 (numpy's axis-0 order).
 *   peak_normalize: every mode.
 *   polar_coordinates: radii, folded (or full) angles and amplitude weights of a stereo pair; weights may be NULL.
 *       Radii and weights are bit-exact, angles within 2 ulp of numpy's arctan2.
 * Workspace: vnd_dsp_workspace(largest element count) bytes. */
int vnd_dsp_workspace(int64_t elems, size_t* bytes);
int vnd_rms_normalize_dev(const vnd_signal* x, int32_t x_ndim, const vnd_signal* y, int32_t y_ndim, int32_t stereo_mode, double epsilon,
                          void* workspace, size_t workspace_bytes, void* stream);
int vnd_peak_normalize_dev(const vnd_signal* y, int32_t ndim, int32_t stereo_mode, double epsilon, void* workspace, size_t workspace_bytes,
                           void* stream);
int vnd_polar_dev(const void* left, const void* right, int64_t n, int32_t dtype, int32_t mode_ms, int32_t semicircular, int32_t normalize,
                  void* radii, void* thetas, void* weights, void* workspace, size_t workspace_bytes, void* stream);

/* Batched stereo-image objective for velvet-noise candidates (optimization.py:46-105 evaluated as
 * optimization.py:108-117 does, for the candidates optimization.py:260-272 builds: channel 0
 * filtered, channel 1 passed through, LR, no normaliser).
 *   clips       : n_clips planar stereo clips, element(clip, c, t) = clips[clip*clip_stride + c*chan_stride + t]
 *   cand        : SEGMENTED tap program with one "channel" per candidate (channels = n_cand)
 *   partials    : n_clips * n_cand * 12 float64 on the device:
 *                 [sum r, sum r*th, sum r*th^2, sum r*th^3, sum L*R, sum L*L,
 *                  |d|, |s| of the frame with the largest |d|/|s| among frames with s >= 0,
 *                  |d|, |s| of the same among frames with s < 0,  frames, 0]
 *                 with d = L - R, s = L + R, th = atan(d / s) (the folded angle, |err| <= 1e-7);
 *                 max|theta| is evaluated by the host from the two tracked frames (rounding is
 *                 monotone, so they are where the reference's float32 maximum is attained).
 * The host combines the partials with the reference's dtype chain. */
int vnd_objective_workspace(int64_t frames, int32_t n_clips, int32_t n_cand, size_t* bytes);
int vnd_vn_objective_batch_dev(const float* clips, int64_t frames, int32_t n_clips, int64_t clip_stride,
                               int64_t chan_stride, const vnd_tap_program* cand, double* partials,
                               void* workspace, size_t workspace_bytes, void* stream);

/* Haas candidates (optimization.py:183-203): LR mode, channel 0 delayed by delays[k] frames,
 * float64 arithmetic, frames + delays[k] output frames.  clips may be F32 or F64 (clip_dtype);
 * with delays = {0} this is the objective of an arbitrary stereo signal.  8 float64 per pair:
 *   [sum r, sum r*th, sum r*th^2, sum r*th^3, max|th|, sum L*R, sum L*L, frames] */
int vnd_haas_objective_batch_dev(const void* clips, int32_t clip_dtype, int64_t frames, int32_t n_clips, int64_t clip_stride,
                                 int64_t chan_stride, const int32_t* delays, int32_t n_cand, double* partials,
                                 void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------ host API */
typedef struct vnd_ctx vnd_ctx;

int vnd_ctx_create(int device, vnd_ctx** ctx);
int vnd_ctx_destroy(vnd_ctx* ctx);
/* Page-locked host memory for callers that want zero-staging transfers: buffers from these calls (or registered
 * with cudaHostRegister) are copied by DMA directly; any other host buffer is staged through the context's own
 * page-locked ring.  vnd_host_alloc allocates on the calling thread's current device, vnd_ctx_host_alloc on the
 * context's device (it never touches another GPU). */
int vnd_host_alloc(size_t bytes, void** ptr);
int vnd_ctx_host_alloc(vnd_ctx* ctx, size_t bytes, void** ptr);
int vnd_host_free(void* ptr);

/* Host-buffer versions of the calls above (all pointers, including taps->words/offsets, are HOST
 * pointers).  Each call uploads, runs and downloads; on return the result is in `y`/`out`.
 * vnd_sparse_fir_host overlaps the three for float32 slabs of 8 MB and more: planar (channels, frames) slabs are
 * cut into channel groups, C-order (frames, channels) slabs - the reference's layout - into runs of frames
 * uploaded with the filter's halo (wide ones are transposed to planar on the device around the kernel), and
 * consecutive chunks flow through a three-deep device ring on three streams. */
int vnd_sparse_fir_host(vnd_ctx* ctx, const vnd_signal* x, const vnd_signal* y, const vnd_tap_program* taps);
int vnd_vn_decorrelate_host(vnd_ctx* ctx, const vnd_signal* x, const vnd_signal* out,
                            const vnd_tap_program* taps, const vnd_epilogue* ep);
int vnd_haas_host(vnd_ctx* ctx, const vnd_signal* x, const vnd_signal* out, int32_t delay,
                  int32_t delayed_channel, int32_t mode_ms, int32_t mono, int32_t use_width, double width);
int vnd_stereo_op_host(vnd_ctx* ctx, const vnd_signal* a, const vnd_signal* dry, int32_t op, double width);
int vnd_rms_normalize_host(vnd_ctx* ctx, const vnd_signal* x, int32_t x_ndim, const vnd_signal* y, int32_t y_ndim, int32_t stereo_mode,
                           double epsilon);
int vnd_peak_normalize_host(vnd_ctx* ctx, const vnd_signal* y, int32_t ndim, int32_t stereo_mode, double epsilon);
int vnd_polar_host(vnd_ctx* ctx, const void* left, const void* right, int64_t n, int32_t dtype, int32_t mode_ms, int32_t semicircular,
                   int32_t normalize, void* radii, void* thetas, void* weights);
int vnd_vn_objective_batch_host(vnd_ctx* ctx, const float* clips, int64_t frames, int32_t n_clips,
                                int64_t clip_stride, int64_t chan_stride, const vnd_tap_program* cand,
                                double* partials);
int vnd_haas_objective_batch_host(vnd_ctx* ctx, const void* clips, int32_t clip_dtype, int64_t frames, int32_t n_clips,
                                  int64_t clip_stride, int64_t chan_stride, const int32_t* delays,
                                  int32_t n_cand, double* partials);

/* The overlapped path of vnd_sparse_fir_host with an explicit stage size: x and y are HOST planar
 * (channels, frames) float32 slabs; channels are cut into groups of `channels_per_chunk`, and
 * upload / kernel / download of consecutive groups overlap on three streams.  Page-locked buffers
 * (vnd_host_alloc) are copied directly; pageable ones go through the context's page-locked staging ring,
 * filled and drained by two helper threads. */
int vnd_sparse_fir_stream_host(vnd_ctx* ctx, const float* x, float* y, int64_t frames, int32_t channels,
                               const vnd_tap_program* taps, int32_t channels_per_chunk);

/* Number of kernels this library has launched in the calling process (all threads); bench.py
 * reports the difference across its timed region as `gpu_launches`. */
int64_t vnd_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* VND_B200_H */
