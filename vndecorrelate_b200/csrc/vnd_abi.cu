// extern "C" surface of libvnd_b200.so (see include/vnd_b200.h): argument checking, the device
// entry points, and the host-buffer entry points with their context (device arena, streams,
// pinned staging).

#include <stdarg.h>
#include <string.h>

#include <mutex>
#include <vector>

#include "vnd_common.cuh"

namespace vnd {

// ---- launchers implemented in the kernel translation units ----------------------------------
int sparse_fir_launch(const vnd_signal* x, const vnd_signal* y, const vnd_tap_program* taps, int max_prog_words, cudaStream_t st);
int vn_stereo_launch(const vnd_signal* x, void* out, int out_dtype, long long o_st, long long o_sc, const vnd_tap_program* taps,
                     int prog_words, const vnd_epilogue* ep, const float* gains, int delay, int delay_ch, cudaStream_t st);
int seq_sumsq_launch(const vnd_signal* a, const vnd_signal* b, void* sums, cudaStream_t st);
int rms_gain_launch(const void* sums, void* gains, int channels, long long frames, int dtype, cudaStream_t st);
int place_launch(const float* y, long long y_st, long long y_sc, long long frames, int channels, const vnd_signal* out,
                 const float* gains, int delay, int delay_ch, cudaStream_t st);
int scale_launch(float* y, long long y_st, long long y_sc, long long frames, int channels, const float* gains, cudaStream_t st);
int haas_launch(const vnd_signal* x, const vnd_signal* out, int delay, int delay_ch, int mode_ms, int mono, int use_width,
                double width, cudaStream_t st);
int stereo_op_launch(const vnd_signal* a, const vnd_signal* dry, int op, double width, const void* gains, cudaStream_t st);
int transpose_launch(const float* src, float* dst, long long rows, long long cols, cudaStream_t st);
int objective_workspace_bytes(long long frames, int n_clips, int n_cand, size_t* bytes);
int vn_objective_launch(const float* clips, long long frames, int n_clips, long long clip_stride, long long chan_stride,
                        const vnd_tap_program* cand, double* partials, void* workspace, size_t workspace_bytes, cudaStream_t st);
int haas_objective_launch(const void* clips, int clip_dtype, long long frames, int n_clips, long long clip_stride, long long chan_stride,
                          const int* delays, int n_cand, double* partials, cudaStream_t st);

// ---- error state ----------------------------------------------------------------------------
static thread_local char t_error[512] = "";
std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(t_error, sizeof(t_error), fmt, ap);
  va_end(ap);
}

int device_info(DeviceInfo* out) {
  static std::mutex mu;
  static DeviceInfo cache[64];
  static bool have[64] = {false};
  int dev = 0;
  VND_CUDA_OK(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lock(mu);
  if (dev < 0 || dev >= 64) dev = 0;
  if (!have[dev]) {
    VND_CUDA_OK(cudaDeviceGetAttribute(&cache[dev].sm_count, cudaDevAttrMultiProcessorCount, dev));
    VND_CUDA_OK(cudaDeviceGetAttribute(&cache[dev].max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    have[dev] = true;
  }
  *out = cache[dev];
  return VND_OK;
}

static int check_signal(const vnd_signal* s, const char* name, bool need_data = true) {
  VND_REQUIRE(s != nullptr, VND_EINVAL, "%s is null", name);
  VND_REQUIRE(s->frames >= 0 && s->channels >= 0, VND_EINVAL, "%s has a negative extent", name);
  VND_REQUIRE(s->dtype == VND_F32 || s->dtype == VND_F64, VND_EINVAL, "%s has an unknown dtype %d", name, s->dtype);
  VND_REQUIRE(!need_data || s->data != nullptr || s->frames * (long long)s->channels == 0, VND_EINVAL, "%s has no data pointer", name);
  return VND_OK;
}

static int check_taps(const vnd_tap_program* t) {
  VND_REQUIRE(t != nullptr, VND_EINVAL, "tap program is null");
  VND_REQUIRE(t->channels >= 0 && t->n_words >= 0, VND_EINVAL, "tap program has a negative extent");
  VND_REQUIRE(t->channels == 0 || t->offsets != nullptr, VND_EPROGRAM, "tap program has no offsets");
  VND_REQUIRE(t->n_words == 0 || t->words != nullptr, VND_EPROGRAM, "tap program has no words");
  VND_REQUIRE(t->order >= VND_ORDER_SEGMENTED && t->order <= VND_ORDER_ASCENDING_F64, VND_EINVAL, "unknown tap order %d", t->order);
  VND_REQUIRE(t->halo >= 0 && t->max_channel_words >= 0, VND_EPROGRAM, "tap program has a negative halo or block size");
  return VND_OK;
}

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace vnd

using namespace vnd;

// ================================================================================================
// housekeeping
// ================================================================================================
extern "C" int vnd_abi_version(void) { return VND_ABI_VERSION; }
extern "C" const char* vnd_version(void) { return "vnd_b200 0.1.0 (sm_100a)"; }
extern "C" const char* vnd_last_error(void) { return t_error; }
extern "C" int64_t vnd_launch_count(void) { return (int64_t)g_launches.load(); }

extern "C" const char* vnd_status_string(int status) {
  switch (status) {
    case VND_OK: return "ok";
    case VND_EINVAL: return "invalid argument";
    case VND_ECUDA: return "CUDA runtime error";
    case VND_EUNSUPPORTED: return "unsupported request";
    case VND_ENOMEM: return "workspace or memory too small";
    case VND_EPROGRAM: return "malformed tap program";
    default: return "unknown status";
  }
}

extern "C" int vnd_device_count(int* count) {
  VND_REQUIRE(count != nullptr, VND_EINVAL, "count is null");
  *count = 0;
  VND_CUDA_OK(cudaGetDeviceCount(count));
  return VND_OK;
}

// ================================================================================================
// device entry points
// ================================================================================================
extern "C" int vnd_sparse_fir_dev(const vnd_signal* x, const vnd_signal* y, const vnd_tap_program* taps, void* stream) {
  int rc;
  if ((rc = check_signal(x, "x")) || (rc = check_signal(y, "y")) || (rc = check_taps(taps))) return rc;
  VND_REQUIRE(y->dtype == VND_F32, VND_EINVAL, "y must be float32 (the reference's output buffers are float32)");
  VND_REQUIRE(y->frames == x->frames, VND_EINVAL, "x has %lld frames but y has %lld", (long long)x->frames, (long long)y->frames);
  VND_REQUIRE(y->channels == taps->channels, VND_EINVAL, "y has %d channels but the tap program has %d", y->channels, taps->channels);
  VND_REQUIRE(x->channels >= taps->channels || x->stride_c == 0, VND_EINVAL, "x has %d channels, fewer than the %d outputs",
              x->channels, taps->channels);
  return sparse_fir_launch(x, y, taps, taps->max_channel_words, (cudaStream_t)stream);
}

extern "C" int vnd_vn_decorrelate_workspace(int64_t frames, int32_t channels, const vnd_epilogue* ep, size_t* bytes) {
  VND_REQUIRE(bytes != nullptr && ep != nullptr, VND_EINVAL, "null argument");
  VND_REQUIRE(frames >= 0 && channels >= 0, VND_EINVAL, "negative extent");
  *bytes = 256 + align_up((size_t)channels * 3 * sizeof(float), 256) + (size_t)frames * channels * sizeof(float);
  return VND_OK;
}

extern "C" int vnd_vn_decorrelate_dev(const vnd_signal* x, const vnd_signal* out, const vnd_tap_program* taps,
                                      const vnd_epilogue* ep, void* workspace, size_t workspace_bytes, void* stream) {
  int rc;
  if ((rc = check_signal(x, "x")) || (rc = check_signal(out, "out")) || (rc = check_taps(taps))) return rc;
  VND_REQUIRE(ep != nullptr, VND_EINVAL, "epilogue is null");
  VND_REQUIRE(x->dtype == VND_F32, VND_EINVAL, "x must be float32 (decorrelate casts first, decorrelation.py:426)");
  VND_REQUIRE(taps->order == VND_ORDER_SEGMENTED, VND_EINVAL, "decorrelate needs a SEGMENTED tap program");
  const int C = taps->channels;
  const long long L = x->frames;
  VND_REQUIRE(ep->haas_delay >= 0, VND_EINVAL, "negative Haas delay");
  VND_REQUIRE(out->channels == C, VND_EINVAL, "out has %d channels but the tap program has %d", out->channels, C);
  VND_REQUIRE(out->frames == L + ep->haas_delay, VND_EINVAL, "out must have frames + haas_delay = %lld frames, has %lld",
              L + ep->haas_delay, (long long)out->frames);
  VND_REQUIRE(x->channels >= C || x->stride_c == 0, VND_EINVAL, "x has fewer channels than the program");
  if (ep->ms_encode || ep->use_width || ep->haas_delay > 0)
    VND_REQUIRE(C == 2, VND_EINVAL, "M/S encode, width and the Haas delay need exactly 2 channels (utils/dsp.py:297-302)");
  if (ep->haas_delay > 0) VND_REQUIRE(ep->haas_channel == 0 || ep->haas_channel == 1, VND_EINVAL, "haas_channel must be 0 or 1");
  cudaStream_t st = (cudaStream_t)stream;
  const int delay = ep->haas_delay, dch = ep->haas_channel;

  size_t need = 0;
  vnd_vn_decorrelate_workspace(L, C, ep, &need);
  char* ws = reinterpret_cast<char*>(workspace);
  float* sums = reinterpret_cast<float*>(ws);
  float* gains = sums + 2 * (size_t)C;
  float* scratch = reinterpret_cast<float*>(ws + align_up((size_t)C * 3 * sizeof(float), 256));

  if (L == 0 || C == 0) return place_launch(nullptr, 0, 0, 0, C, out, nullptr, delay, dch, st);

  // one kernel, one read and one write per sample
  if (!ep->rms_normalize && C == 2) {
    rc = vn_stereo_launch(x, out->data, out->dtype, out->stride_t, out->stride_c, taps, taps->max_channel_words * 2, ep, nullptr,
                          delay, dch, st);
    if (rc != VND_EUNSUPPORTED) return rc;
  }
  VND_REQUIRE(workspace != nullptr && workspace_bytes >= need, VND_ENOMEM, "workspace too small: need %zu bytes, have %zu", need,
              workspace_bytes);
  vnd_signal y{scratch, L, C, VND_F32, C, 1};
  rc = VND_EUNSUPPORTED;
  if (C == 2) rc = vn_stereo_launch(x, scratch, VND_F32, 2, 1, taps, taps->max_channel_words * 2, ep, nullptr, 0, 0, st);
  if (rc == VND_EUNSUPPORTED) {  // unfused composition (halo too long for the stereo kernel, or C != 2)
    if ((rc = sparse_fir_launch(x, &y, taps, taps->max_channel_words, st))) return rc;
    if (ep->ms_encode && (rc = stereo_op_launch(&y, x, 3, 0.0, nullptr, st))) return rc;
    if (ep->use_width && (rc = stereo_op_launch(&y, nullptr, 2, ep->width, nullptr, st))) return rc;
  } else if (rc) {
    return rc;
  }
  const float* g = nullptr;
  if (ep->rms_normalize) {
    vnd_signal xa = *x;
    xa.channels = C;
    if ((rc = seq_sumsq_launch(&xa, &y, sums, st))) return rc;
    if ((rc = rms_gain_launch(sums, gains, C, L, VND_F32, st))) return rc;
    g = gains;
  }
  return place_launch(scratch, C, 1, L, C, out, g, delay, dch, st);
}

extern "C" int vnd_colsumsq_seq_f32_dev(const vnd_signal* a, float* sums, void* stream) {
  int rc;
  if ((rc = check_signal(a, "a"))) return rc;
  VND_REQUIRE(a->dtype == VND_F32, VND_EINVAL, "a must be float32");
  VND_REQUIRE(sums != nullptr || a->channels == 0, VND_EINVAL, "sums is null");
  return seq_sumsq_launch(a, nullptr, sums, (cudaStream_t)stream);
}

extern "C" int vnd_haas_dev(const vnd_signal* x, const vnd_signal* out, int32_t delay, int32_t delayed_channel, int32_t mode_ms,
                            int32_t mono, int32_t use_width, double width, void* stream) {
  int rc;
  if ((rc = check_signal(x, "x")) || (rc = check_signal(out, "out"))) return rc;
  VND_REQUIRE(x->dtype == VND_F32 && out->dtype == VND_F64, VND_EINVAL, "Haas takes float32 in and float64 out (decorrelation.py:194,206)");
  VND_REQUIRE(delay >= 0, VND_EINVAL, "negative delay");
  VND_REQUIRE(delayed_channel == 0 || delayed_channel == 1, VND_EINVAL, "delayed_channel must be 0 or 1");
  VND_REQUIRE(out->channels == 2 && out->frames == x->frames + delay, VND_EINVAL, "out must be (frames + delay, 2)");
  return haas_launch(x, out, delay, delayed_channel, mode_ms, mono, use_width, width, (cudaStream_t)stream);
}

extern "C" int vnd_stereo_op_dev(const vnd_signal* a, const vnd_signal* dry, int32_t op, double width, void* workspace,
                                 size_t workspace_bytes, void* stream) {
  int rc;
  if ((rc = check_signal(a, "a"))) return rc;
  VND_REQUIRE(op >= 0 && op <= 4, VND_EINVAL, "unknown stereo op %d", op);
  VND_REQUIRE(a->channels == 2, VND_EINVAL, "stereo helpers need a (frames, 2) signal (utils/dsp.py:297-302)");
  cudaStream_t st = (cudaStream_t)stream;
  if (op == 3 || op == 4) {
    if ((rc = check_signal(dry, "dry"))) return rc;
    VND_REQUIRE(dry->channels == 2 && dry->frames == a->frames && dry->dtype == a->dtype, VND_EINVAL,
                "dry must match a in shape and dtype");
  }
  if (op == 4) {
    const size_t esz = a->dtype == VND_F64 ? 8 : 4;
    VND_REQUIRE(workspace != nullptr && workspace_bytes >= 6 * esz, VND_ENOMEM, "rms workspace needs %zu bytes", 6 * esz);
    char* sums = reinterpret_cast<char*>(workspace);
    char* gains = sums + 4 * esz;
    if ((rc = seq_sumsq_launch(dry, a, sums, st))) return rc;
    if ((rc = rms_gain_launch(sums, gains, 2, a->frames, a->dtype, st))) return rc;
    return stereo_op_launch(a, nullptr, 4, 0.0, gains, st);
  }
  return stereo_op_launch(a, op == 3 ? dry : nullptr, op, width, nullptr, st);
}

extern "C" int vnd_objective_workspace(int64_t frames, int32_t n_clips, int32_t n_cand, size_t* bytes) {
  VND_REQUIRE(bytes != nullptr, VND_EINVAL, "bytes is null");
  VND_REQUIRE(frames >= 0 && n_clips >= 0 && n_cand >= 0, VND_EINVAL, "negative extent");
  return objective_workspace_bytes(frames, n_clips, n_cand, bytes);
}

extern "C" int vnd_vn_objective_batch_dev(const float* clips, int64_t frames, int32_t n_clips, int64_t clip_stride, int64_t chan_stride,
                                          const vnd_tap_program* cand, double* partials, void* workspace, size_t workspace_bytes,
                                          void* stream) {
  int rc;
  if ((rc = check_taps(cand))) return rc;
  VND_REQUIRE(cand->order == VND_ORDER_SEGMENTED, VND_EINVAL, "objective candidates need SEGMENTED tap programs");
  VND_REQUIRE(frames >= 0 && n_clips >= 0, VND_EINVAL, "negative extent");
  VND_REQUIRE((clips != nullptr && partials != nullptr) || n_clips == 0 || cand->channels == 0, VND_EINVAL, "null buffer");
  return vn_objective_launch(clips, frames, n_clips, clip_stride, chan_stride, cand, partials, workspace, workspace_bytes,
                             (cudaStream_t)stream);
}

extern "C" int vnd_haas_objective_batch_dev(const void* clips, int32_t clip_dtype, int64_t frames, int32_t n_clips, int64_t clip_stride,
                                            int64_t chan_stride, const int32_t* delays, int32_t n_cand, double* partials, void* workspace,
                                            size_t workspace_bytes, void* stream) {
  (void)workspace;
  (void)workspace_bytes;
  VND_REQUIRE(frames >= 0 && n_clips >= 0 && n_cand >= 0, VND_EINVAL, "negative extent");
  VND_REQUIRE((clips != nullptr && partials != nullptr && delays != nullptr) || n_clips == 0 || n_cand == 0, VND_EINVAL, "null buffer");
  VND_REQUIRE(clip_dtype == VND_F32 || clip_dtype == VND_F64, VND_EINVAL, "unknown clip dtype %d", clip_dtype);
  return haas_objective_launch(clips, clip_dtype, frames, n_clips, clip_stride, chan_stride, delays, n_cand, partials, (cudaStream_t)stream);
}

// ================================================================================================
// host entry points
// ================================================================================================
struct vnd_ctx {
  int device = 0;
  cudaStream_t streams[3] = {nullptr, nullptr, nullptr};
  struct Buf {
    void* p = nullptr;
    size_t cap = 0;
  };
  Buf slots[12];  // grow-only device arena
  std::mutex mu;
};

namespace {

enum Slot { S_IN = 0, S_OUT = 1, S_WORK = 2, S_WORDS = 3, S_OFFS = 4, S_AUX = 5, S_RING = 6 /* 6..11 */ };

int arena(vnd_ctx* ctx, int slot, size_t bytes, void** out) {
  vnd_ctx::Buf& b = ctx->slots[slot];
  if (bytes == 0) bytes = 16;
  if (b.cap < bytes) {
    if (b.p) VND_CUDA_OK(cudaFree(b.p));
    b.p = nullptr;
    b.cap = 0;
    const size_t want = align_up(bytes + bytes / 8, 1 << 20);
    cudaError_t e = cudaMalloc(&b.p, want);
    if (e != cudaSuccess) {
      (void)cudaGetLastError();
      set_error("device arena: cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e));
      return VND_ENOMEM;
    }
    b.cap = want;
  }
  *out = b.p;
  return VND_OK;
}

// A strided host view is "dense" when it covers one contiguous block: C-order (frames, channels),
// planar (channels, frames), or a mono signal broadcast with stride_c == 0.
struct Dense {
  size_t elems;
  bool ok;
};
Dense dense_extent(const vnd_signal* s) {
  const long long L = s->frames, C = s->channels;
  if (L * C == 0) return {0, true};
  if (s->stride_c == 0 && s->stride_t == 1) return {(size_t)L, true};
  if (C == 1 && s->stride_t == 1) return {(size_t)L, true};
  if (s->stride_c == 1 && s->stride_t == C) return {(size_t)(L * C), true};
  if (s->stride_t == 1 && s->stride_c == L) return {(size_t)(L * C), true};
  return {0, false};
}

size_t esize(int dtype) { return dtype == VND_F64 ? 8 : 4; }

int upload_program(vnd_ctx* ctx, const vnd_tap_program* host, vnd_tap_program* dev, cudaStream_t st) {
  *dev = *host;
  void *w = nullptr, *o = nullptr;
  int rc;
  if ((rc = arena(ctx, S_WORDS, (size_t)host->n_words * 4, &w))) return rc;
  if ((rc = arena(ctx, S_OFFS, ((size_t)host->channels + 1) * 4, &o))) return rc;
  if (host->n_words) VND_CUDA_OK(cudaMemcpyAsync(w, host->words, (size_t)host->n_words * 4, cudaMemcpyHostToDevice, st));
  if (host->channels) VND_CUDA_OK(cudaMemcpyAsync(o, host->offsets, ((size_t)host->channels + 1) * 4, cudaMemcpyHostToDevice, st));
  dev->words = reinterpret_cast<const int32_t*>(w);
  dev->offsets = reinterpret_cast<const int32_t*>(o);
  return VND_OK;
}

struct DeviceGuard {
  int prev = -1;
  bool ok = true;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    ok = cudaSetDevice(dev) == cudaSuccess;
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

#define VND_ENTER(ctx)                                                      \
  VND_REQUIRE((ctx) != nullptr, VND_EINVAL, "context is null");             \
  std::lock_guard<std::mutex> _lock((ctx)->mu);                             \
  DeviceGuard _guard((ctx)->device);                                        \
  VND_REQUIRE(_guard.ok, VND_ECUDA, "cudaSetDevice(%d) failed", (ctx)->device)

}  // namespace

extern "C" int vnd_ctx_create(int device, vnd_ctx** out) {
  VND_REQUIRE(out != nullptr, VND_EINVAL, "ctx out pointer is null");
  *out = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    (void)cudaGetLastError();
    set_error("no CUDA device available (%s); this library has no CPU fallback", e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    return VND_ECUDA;
  }
  VND_REQUIRE(device >= 0 && device < n, VND_EINVAL, "device %d out of range (have %d)", device, n);
  vnd_ctx* ctx = new vnd_ctx();
  ctx->device = device;
  DeviceGuard guard(device);
  for (auto& s : ctx->streams) {
    if (cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking) != cudaSuccess) {
      set_error("cudaStreamCreate failed");
      delete ctx;
      return VND_ECUDA;
    }
  }
  *out = ctx;
  return VND_OK;
}

extern "C" int vnd_ctx_destroy(vnd_ctx* ctx) {
  if (!ctx) return VND_OK;
  {
    DeviceGuard guard(ctx->device);
    for (auto& s : ctx->streams)
      if (s) {
        cudaStreamSynchronize(s);
        cudaStreamDestroy(s);
      }
    for (auto& b : ctx->slots)
      if (b.p) cudaFree(b.p);
  }
  delete ctx;
  return VND_OK;
}

extern "C" int vnd_host_alloc(size_t bytes, void** ptr) {
  VND_REQUIRE(ptr != nullptr, VND_EINVAL, "ptr is null");
  *ptr = nullptr;
  VND_CUDA_OK(cudaHostAlloc(ptr, bytes ? bytes : 16, cudaHostAllocPortable));
  return VND_OK;
}

extern "C" int vnd_host_free(void* ptr) {
  if (ptr) VND_CUDA_OK(cudaFreeHost(ptr));
  return VND_OK;
}

extern "C" int vnd_sparse_fir_host(vnd_ctx* ctx, const vnd_signal* x, const vnd_signal* y, const vnd_tap_program* taps) {
  VND_ENTER(ctx);
  int rc;
  if ((rc = check_signal(x, "x")) || (rc = check_signal(y, "y")) || (rc = check_taps(taps))) return rc;
  const Dense dx = dense_extent(x), dy = dense_extent(y);
  VND_REQUIRE(dx.ok && dy.ok, VND_EUNSUPPORTED, "host signals must be contiguous C-order (frames, channels) or planar (channels, frames)");
  cudaStream_t st = ctx->streams[0];
  vnd_tap_program dprog;
  if ((rc = upload_program(ctx, taps, &dprog, st))) return rc;
  void *din = nullptr, *dout = nullptr;
  if ((rc = arena(ctx, S_IN, dx.elems * esize(x->dtype), &din))) return rc;
  if ((rc = arena(ctx, S_OUT, dy.elems * 4, &dout))) return rc;
  if (dx.elems) VND_CUDA_OK(cudaMemcpyAsync(din, x->data, dx.elems * esize(x->dtype), cudaMemcpyHostToDevice, st));
  vnd_signal xd = *x, yd = *y;
  xd.data = din;
  yd.data = dout;
  // Wide frame-interleaved float32 slabs: transpose to planar on the device so that the tile loads
  // and stores are coalesced, run the planar kernel, transpose back.
  const bool wide_interleaved = x->dtype == VND_F32 && taps->channels > 2 && x->stride_c == 1 && x->stride_t == x->channels &&
                                x->channels == taps->channels && y->stride_c == 1 && y->stride_t == y->channels && x->frames > 0;
  if (wide_interleaved) {
    void *pin = nullptr, *pout = nullptr;
    if ((rc = arena(ctx, S_WORK, dx.elems * 4, &pin))) return rc;
    if ((rc = arena(ctx, S_AUX, dy.elems * 4, &pout))) return rc;
    if ((rc = transpose_launch((const float*)din, (float*)pin, x->frames, x->channels, st))) return rc;
    vnd_signal xp{pin, x->frames, x->channels, VND_F32, 1, x->frames};
    vnd_signal yp{pout, y->frames, y->channels, VND_F32, 1, y->frames};
    if ((rc = vnd_sparse_fir_dev(&xp, &yp, &dprog, st))) return rc;
    if ((rc = transpose_launch((const float*)pout, (float*)dout, y->channels, y->frames, st))) return rc;
  } else {
    if ((rc = vnd_sparse_fir_dev(&xd, &yd, &dprog, st))) return rc;
  }
  if (dy.elems) VND_CUDA_OK(cudaMemcpyAsync(y->data, dout, dy.elems * 4, cudaMemcpyDeviceToHost, st));
  VND_CUDA_OK(cudaStreamSynchronize(st));
  return VND_OK;
}

extern "C" int vnd_vn_decorrelate_host(vnd_ctx* ctx, const vnd_signal* x, const vnd_signal* out, const vnd_tap_program* taps,
                                       const vnd_epilogue* ep) {
  VND_ENTER(ctx);
  int rc;
  if ((rc = check_signal(x, "x")) || (rc = check_signal(out, "out")) || (rc = check_taps(taps))) return rc;
  VND_REQUIRE(ep != nullptr, VND_EINVAL, "epilogue is null");
  const Dense dx = dense_extent(x), dy = dense_extent(out);
  VND_REQUIRE(dx.ok && dy.ok, VND_EUNSUPPORTED, "host signals must be contiguous C-order (frames, channels) or planar (channels, frames)");
  cudaStream_t st = ctx->streams[0];
  vnd_tap_program dprog;
  if ((rc = upload_program(ctx, taps, &dprog, st))) return rc;
  void *din = nullptr, *dout = nullptr, *work = nullptr;
  size_t wbytes = 0;
  if ((rc = vnd_vn_decorrelate_workspace(x->frames, taps->channels, ep, &wbytes))) return rc;
  if ((rc = arena(ctx, S_IN, dx.elems * esize(x->dtype), &din))) return rc;
  if ((rc = arena(ctx, S_OUT, dy.elems * esize(out->dtype), &dout))) return rc;
  if ((rc = arena(ctx, S_WORK, wbytes, &work))) return rc;
  if (dx.elems) VND_CUDA_OK(cudaMemcpyAsync(din, x->data, dx.elems * esize(x->dtype), cudaMemcpyHostToDevice, st));
  vnd_signal xd = *x, od = *out;
  xd.data = din;
  od.data = dout;
  if ((rc = vnd_vn_decorrelate_dev(&xd, &od, &dprog, ep, work, wbytes, st))) return rc;
  if (dy.elems) VND_CUDA_OK(cudaMemcpyAsync(out->data, dout, dy.elems * esize(out->dtype), cudaMemcpyDeviceToHost, st));
  VND_CUDA_OK(cudaStreamSynchronize(st));
  return VND_OK;
}

extern "C" int vnd_haas_host(vnd_ctx* ctx, const vnd_signal* x, const vnd_signal* out, int32_t delay, int32_t delayed_channel,
                             int32_t mode_ms, int32_t mono, int32_t use_width, double width) {
  VND_ENTER(ctx);
  int rc;
  if ((rc = check_signal(x, "x")) || (rc = check_signal(out, "out"))) return rc;
  const Dense dx = dense_extent(x), dy = dense_extent(out);
  VND_REQUIRE(dx.ok && dy.ok, VND_EUNSUPPORTED, "host signals must be contiguous");
  cudaStream_t st = ctx->streams[0];
  void *din = nullptr, *dout = nullptr;
  if ((rc = arena(ctx, S_IN, dx.elems * esize(x->dtype), &din))) return rc;
  if ((rc = arena(ctx, S_OUT, dy.elems * esize(out->dtype), &dout))) return rc;
  if (dx.elems) VND_CUDA_OK(cudaMemcpyAsync(din, x->data, dx.elems * esize(x->dtype), cudaMemcpyHostToDevice, st));
  vnd_signal xd = *x, od = *out;
  xd.data = din;
  od.data = dout;
  if ((rc = vnd_haas_dev(&xd, &od, delay, delayed_channel, mode_ms, mono, use_width, width, st))) return rc;
  if (dy.elems) VND_CUDA_OK(cudaMemcpyAsync(out->data, dout, dy.elems * esize(out->dtype), cudaMemcpyDeviceToHost, st));
  VND_CUDA_OK(cudaStreamSynchronize(st));
  return VND_OK;
}

extern "C" int vnd_stereo_op_host(vnd_ctx* ctx, const vnd_signal* a, const vnd_signal* dry, int32_t op, double width) {
  VND_ENTER(ctx);
  int rc;
  if ((rc = check_signal(a, "a"))) return rc;
  const Dense da = dense_extent(a);
  VND_REQUIRE(da.ok, VND_EUNSUPPORTED, "host signals must be contiguous");
  cudaStream_t st = ctx->streams[0];
  void *dA = nullptr, *dD = nullptr, *work = nullptr;
  if ((rc = arena(ctx, S_IN, da.elems * esize(a->dtype), &dA))) return rc;
  if ((rc = arena(ctx, S_WORK, 256, &work))) return rc;
  if (da.elems) VND_CUDA_OK(cudaMemcpyAsync(dA, a->data, da.elems * esize(a->dtype), cudaMemcpyHostToDevice, st));
  vnd_signal ad = *a, dd{};
  ad.data = dA;
  if (op == 3 || op == 4) {
    if ((rc = check_signal(dry, "dry"))) return rc;
    const Dense d2 = dense_extent(dry);
    VND_REQUIRE(d2.ok, VND_EUNSUPPORTED, "host signals must be contiguous");
    if ((rc = arena(ctx, S_OUT, d2.elems * esize(dry->dtype), &dD))) return rc;
    if (d2.elems) VND_CUDA_OK(cudaMemcpyAsync(dD, dry->data, d2.elems * esize(dry->dtype), cudaMemcpyHostToDevice, st));
    dd = *dry;
    dd.data = dD;
  }
  if ((rc = vnd_stereo_op_dev(&ad, (op == 3 || op == 4) ? &dd : nullptr, op, width, work, 256, st))) return rc;
  if (da.elems) VND_CUDA_OK(cudaMemcpyAsync(a->data, dA, da.elems * esize(a->dtype), cudaMemcpyDeviceToHost, st));
  VND_CUDA_OK(cudaStreamSynchronize(st));
  return VND_OK;
}

extern "C" int vnd_vn_objective_batch_host(vnd_ctx* ctx, const float* clips, int64_t frames, int32_t n_clips, int64_t clip_stride,
                                           int64_t chan_stride, const vnd_tap_program* cand, double* partials) {
  VND_ENTER(ctx);
  int rc;
  if ((rc = check_taps(cand))) return rc;
  VND_REQUIRE(frames >= 0 && n_clips >= 0, VND_EINVAL, "negative extent");
  if (n_clips == 0 || cand->channels == 0) return VND_OK;
  VND_REQUIRE(clips != nullptr && partials != nullptr, VND_EINVAL, "null buffer");
  VND_REQUIRE(chan_stride >= frames && clip_stride >= chan_stride + frames, VND_EUNSUPPORTED, "clips must be planar: clip, channel, frame");
  cudaStream_t st = ctx->streams[0];
  vnd_tap_program dprog;
  if ((rc = upload_program(ctx, cand, &dprog, st))) return rc;
  const size_t in_bytes = ((size_t)(n_clips - 1) * clip_stride + chan_stride + frames) * 4;
  const size_t out_bytes = (size_t)n_clips * cand->channels * 12 * 8;
  size_t wbytes = 0;
  if ((rc = objective_workspace_bytes(frames, n_clips, cand->channels, &wbytes))) return rc;
  void *din = nullptr, *dout = nullptr, *work = nullptr;
  if ((rc = arena(ctx, S_IN, in_bytes, &din))) return rc;
  if ((rc = arena(ctx, S_OUT, out_bytes, &dout))) return rc;
  if ((rc = arena(ctx, S_WORK, wbytes, &work))) return rc;
  VND_CUDA_OK(cudaMemcpyAsync(din, clips, in_bytes, cudaMemcpyHostToDevice, st));
  if ((rc = vn_objective_launch((const float*)din, frames, n_clips, clip_stride, chan_stride, &dprog, (double*)dout, work, wbytes, st)))
    return rc;
  VND_CUDA_OK(cudaMemcpyAsync(partials, dout, out_bytes, cudaMemcpyDeviceToHost, st));
  VND_CUDA_OK(cudaStreamSynchronize(st));
  return VND_OK;
}

extern "C" int vnd_haas_objective_batch_host(vnd_ctx* ctx, const void* clips, int32_t clip_dtype, int64_t frames, int32_t n_clips, int64_t clip_stride,
                                             int64_t chan_stride, const int32_t* delays, int32_t n_cand, double* partials) {
  VND_ENTER(ctx);
  int rc;
  VND_REQUIRE(frames >= 0 && n_clips >= 0 && n_cand >= 0, VND_EINVAL, "negative extent");
  if (n_clips == 0 || n_cand == 0) return VND_OK;
  VND_REQUIRE(clips != nullptr && partials != nullptr && delays != nullptr, VND_EINVAL, "null buffer");
  VND_REQUIRE(clip_dtype == VND_F32 || clip_dtype == VND_F64, VND_EINVAL, "unknown clip dtype %d", clip_dtype);
  VND_REQUIRE(chan_stride >= frames && clip_stride >= chan_stride + frames, VND_EUNSUPPORTED, "clips must be planar: clip, channel, frame");
  cudaStream_t st = ctx->streams[0];
  const size_t in_bytes = ((size_t)(n_clips - 1) * clip_stride + chan_stride + frames) * esize(clip_dtype);
  const size_t out_bytes = (size_t)n_clips * n_cand * 8 * 8;
  void *din = nullptr, *dout = nullptr, *dd = nullptr;
  if ((rc = arena(ctx, S_IN, in_bytes, &din))) return rc;
  if ((rc = arena(ctx, S_OUT, out_bytes, &dout))) return rc;
  if ((rc = arena(ctx, S_AUX, (size_t)n_cand * 4, &dd))) return rc;
  VND_CUDA_OK(cudaMemcpyAsync(din, clips, in_bytes, cudaMemcpyHostToDevice, st));
  VND_CUDA_OK(cudaMemcpyAsync(dd, delays, (size_t)n_cand * 4, cudaMemcpyHostToDevice, st));
  if ((rc = haas_objective_launch(din, clip_dtype, frames, n_clips, clip_stride, chan_stride, (const int*)dd, n_cand, (double*)dout, st)))
    return rc;
  VND_CUDA_OK(cudaMemcpyAsync(partials, dout, out_bytes, cudaMemcpyDeviceToHost, st));
  VND_CUDA_OK(cudaStreamSynchronize(st));
  return VND_OK;
}

// Streaming planar FIR: channel groups flow through a 3-deep ring of (in, out) device buffers, one
// stream per ring slot, so the upload of group g+1 and the download of group g-1 overlap the
// kernel of group g (PCIe is full duplex).
extern "C" int vnd_sparse_fir_stream_host(vnd_ctx* ctx, const float* x, float* y, int64_t frames, int32_t channels,
                                          const vnd_tap_program* taps, int32_t channels_per_chunk) {
  VND_ENTER(ctx);
  int rc;
  if ((rc = check_taps(taps))) return rc;
  VND_REQUIRE(frames >= 0 && channels >= 0, VND_EINVAL, "negative extent");
  VND_REQUIRE(taps->channels == channels, VND_EINVAL, "tap program has %d channels, slab has %d", taps->channels, channels);
  if (frames == 0 || channels == 0) return VND_OK;
  VND_REQUIRE(x != nullptr && y != nullptr, VND_EINVAL, "null slab");
  if (channels_per_chunk <= 0) channels_per_chunk = 1;
  if (channels_per_chunk > channels) channels_per_chunk = channels;
  vnd_tap_program dprog;
  if ((rc = upload_program(ctx, taps, &dprog, ctx->streams[0]))) return rc;
  VND_CUDA_OK(cudaStreamSynchronize(ctx->streams[0]));
  const size_t chunk_bytes = (size_t)channels_per_chunk * frames * 4;
  void *din[3], *dout[3];
  for (int s = 0; s < 3; ++s) {
    if ((rc = arena(ctx, S_RING + 2 * s, chunk_bytes, &din[s]))) return rc;
    if ((rc = arena(ctx, S_RING + 2 * s + 1, chunk_bytes, &dout[s]))) return rc;
  }
  int g = 0;
  for (int c0 = 0; c0 < channels; c0 += channels_per_chunk, ++g) {
    const int nc = channels - c0 < channels_per_chunk ? channels - c0 : channels_per_chunk;
    const int s = g % 3;
    cudaStream_t st = ctx->streams[s];
    const size_t bytes = (size_t)nc * frames * 4;
    VND_CUDA_OK(cudaMemcpyAsync(din[s], x + (size_t)c0 * frames, bytes, cudaMemcpyHostToDevice, st));
    vnd_signal xd{din[s], frames, nc, VND_F32, 1, frames};
    vnd_signal yd{dout[s], frames, nc, VND_F32, 1, frames};
    vnd_tap_program sub = dprog;
    sub.channels = nc;
    sub.offsets = dprog.offsets + c0;
    if ((rc = vnd_sparse_fir_dev(&xd, &yd, &sub, st))) return rc;
    VND_CUDA_OK(cudaMemcpyAsync(y + (size_t)c0 * frames, dout[s], bytes, cudaMemcpyDeviceToHost, st));
  }
  for (auto& st : ctx->streams) VND_CUDA_OK(cudaStreamSynchronize(st));
  return VND_OK;
}
