// Sparse velvet-noise FIR kernels (sm_100a) and their launchers.
//
// Reference arithmetic reproduced here (paths relative to the reference root):
//   VelvetNoise.convolve          src/vndecorrelate/decorrelation.py:393-415   (SEGMENTED order)
//   convolve_velvet_noise         src/vndecorrelate/decorrelation.py:630-660   (ASCENDING order)
//   VelvetNoise.decorrelate       src/vndecorrelate/decorrelation.py:417-442   (fused epilogue)
//   encode_signal_to_side_channel src/vndecorrelate/utils/dsp.py:40-63
//   apply_stereo_width            src/vndecorrelate/utils/dsp.py:21-37
//
// Design: one CTA owns `tile` consecutive outputs of one channel (or of one stereo pair).  The
// input tile plus a right halo of `halo` samples (the filter is anti-causal: y[t] needs
// x[t + i]) is staged in shared memory — by one TMA bulk copy (cp.async.bulk, UBLKCP) when the
// channel is contiguous in time — together with the channel's tap program.  Lane l of a warp owns
// outputs base + l + 32 r, r < R: for a warp-uniform tap offset the 32 lanes read 32 consecutive
// words (conflict-free LDS), and each of the R stores is one full 128-byte line.  Samples past the
// end of the signal are staged as +0: adding +0 is bit-identical to the reference dropping the tap.

#include <stdlib.h>

#include "vnd_common.cuh"
#include "vnd_fir.cuh"

namespace vnd {

template <typename T>
struct is_f64 { static constexpr bool value = false; };
template <>
struct is_f64<double> { static constexpr bool value = true; };

// One accumulation step in the reference's arithmetic: acc (-|+)= x with fp32 storage.
template <typename TIn>
__device__ __forceinline__ float step_sub(float acc, TIn x) {
  if constexpr (is_f64<TIn>::value) return __double2float_rn(dsub((double)acc, x));
  else return fsub(acc, x);
}
template <typename TIn>
__device__ __forceinline__ float step_add(float acc, TIn x) {
  if constexpr (is_f64<TIn>::value) return __double2float_rn(dadd((double)acc, x));
  else return fadd(acc, x);
}

// Packed float32 arithmetic (sm_100+: FADD2 / FMUL2, two IEEE round-to-nearest operations per instruction on an aligned
// register pair): the same bits as two scalar operations in half the issue slots.  The tile kernel issues one LDS and
// one add per (output, tap) pair; with packed adds the long-filter configuration (300 taps, shared-memory pipe bound)
// spends 1.5 instead of 2 issue slots per pair.
typedef unsigned long long fir_pair_t;
#define VND_FIR_PACKED(name, op)                                                   \
  __device__ __forceinline__ void name(float& a0, float& a1, float b0, float b1) { \
    fir_pair_t ra, rb;                                                             \
    asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(a0), "f"(a1));                   \
    asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(b0), "f"(b1));                   \
    asm(op ".rn.f32x2 %0, %0, %1;" : "+l"(ra) : "l"(rb));                         \
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a0), "=f"(a1) : "l"(ra));                  \
  }
VND_FIR_PACKED(fir_add2, "add")
VND_FIR_PACKED(fir_sub2, "sub")
VND_FIR_PACKED(fir_mul2, "mul")

// Runs one channel's tap program for R outputs per lane.  `px` points at the staged sample of
// this lane's first output; consecutive outputs of the lane are `STRIDE` samples apart.
template <typename TIn, int MODE, int R, int STRIDE>
__device__ __forceinline__ void run_program(const TIn* __restrict__ px, const int* __restrict__ sprog, int apply_gain,
                                            float (&yv)[R]) {
#pragma unroll
  for (int r = 0; r < R; ++r) yv[r] = 0.0f;
  if constexpr (MODE == MODE_SEG) {
    const int S = sprog[0];
    const int* seg = sprog + 1;
    const int* tp = sprog + 1 + 3 * S;
    for (int s = 0; s < S; ++s) {
      const int n_neg = seg[3 * s + 0];
      const int n_pos = seg[3 * s + 1];
      const float gain = __int_as_float(seg[3 * s + 2]);
      float acc[R];
#pragma unroll
      for (int r = 0; r < R; ++r) acc[r] = 0.0f;
      if constexpr (!is_f64<TIn>::value && (R % 2 == 0)) {  // float32: packed adds, the loads of a tap issued together
        for (int k = 0; k < n_neg; ++k) {
          const TIn* q = px + tp[k];
          float t[R];
#pragma unroll
          for (int r = 0; r < R; ++r) t[r] = q[r * STRIDE];
#pragma unroll
          for (int r = 0; r < R; r += 2) fir_sub2(acc[r], acc[r + 1], t[r], t[r + 1]);
        }
        tp += n_neg;
        for (int k = 0; k < n_pos; ++k) {
          const TIn* q = px + tp[k];
          float t[R];
#pragma unroll
          for (int r = 0; r < R; ++r) t[r] = q[r * STRIDE];
#pragma unroll
          for (int r = 0; r < R; r += 2) fir_add2(acc[r], acc[r + 1], t[r], t[r + 1]);
        }
        tp += n_pos;
        if (apply_gain) {
#pragma unroll
          for (int r = 0; r < R; r += 2) fir_mul2(acc[r], acc[r + 1], gain, gain);
        }
#pragma unroll
        for (int r = 0; r < R; r += 2) fir_add2(yv[r], yv[r + 1], acc[r], acc[r + 1]);
      } else {
        for (int k = 0; k < n_neg; ++k) {
          const TIn* q = px + tp[k];
#pragma unroll
          for (int r = 0; r < R; ++r) acc[r] = step_sub<TIn>(acc[r], q[r * STRIDE]);
        }
        tp += n_neg;
        for (int k = 0; k < n_pos; ++k) {
          const TIn* q = px + tp[k];
#pragma unroll
          for (int r = 0; r < R; ++r) acc[r] = step_add<TIn>(acc[r], q[r * STRIDE]);
        }
        tp += n_pos;
        if (apply_gain) {
#pragma unroll
          for (int r = 0; r < R; ++r) acc[r] = fmul(acc[r], gain);
        }
#pragma unroll
        for (int r = 0; r < R; ++r) yv[r] = fadd(yv[r], acc[r]);
      }
    }
  } else {
    constexpr bool kF64 = is_f64<TIn>::value || MODE == MODE_ASC64;
    const int K = sprog[0];
    const int* e = sprog + 1;
    for (int k = 0; k < K; ++k) {
      const TIn* q = px + e[3 * k];
      const double coef = __hiloint2double(e[3 * k + 2], e[3 * k + 1]);
      if constexpr (kF64) {
#pragma unroll
        for (int r = 0; r < R; ++r) yv[r] = __double2float_rn(dadd((double)yv[r], dmul((double)q[r * STRIDE], coef)));
      } else {
        const float cf = (float)coef;  // exact: the coefficient was a float32
#pragma unroll
        for (int r = 0; r < R; ++r) yv[r] = fadd(yv[r], fmul((float)q[r * STRIDE], cf));
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Planar / strided multichannel FIR: one channel x one tile per CTA.
// Shared memory: [0,16) mbarrier | TIn tile[tile + halo] (16-byte aligned) | int program[]
// ------------------------------------------------------------------------------------------------
template <typename TIn, int MODE, int NT, int R>
__global__ void __launch_bounds__(NT) fir_tile_kernel(const FirParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
  TIn* sx = reinterpret_cast<TIn*>(smem_raw + 16);
  const int span = p.tile + p.halo;
  int* sprog = reinterpret_cast<int*>(smem_raw + 16 + (((size_t)span * sizeof(TIn) + 15) & ~(size_t)15));

  const int tid = threadIdx.x;
  const int c = blockIdx.x / p.tiles_per_channel;
  const long long t0 = (long long)(blockIdx.x % p.tiles_per_channel) * p.tile;
  const long long remain = p.frames - t0;
  const int w0 = p.offsets[c];
  const int nprog = p.offsets[c + 1] - w0;
  const TIn* __restrict__ xc = reinterpret_cast<const TIn*>(p.x) + (long long)c * p.x_sc;
  float* __restrict__ yc = p.y + (long long)c * p.y_sc;

  if (nprog == 0) {  // unfiltered channel: copy through (decorrelation.py:399-400)
    const int n = (int)(remain < p.tile ? remain : p.tile);
    for (int i = tid; i < n; i += NT) yc[(t0 + i) * p.y_st] = (float)xc[(t0 + i) * p.x_st];
    return;
  }

  const int nvalid = (int)(remain < span ? remain : span);
  int nbulk = 0;
  if constexpr (!is_f64<TIn>::value) {
    if (p.bulk_ok) {
      nbulk = nvalid & ~3;
      if (tid == 0) {
        mbar_init(bar, 1);
        mbar_fence_init();
      }
      __syncthreads();
      if (tid == 0 && nbulk > 0) {
        mbar_expect_tx(bar, (uint32_t)nbulk * 4u);
        bulk_g2s(sx, xc + t0, (uint32_t)nbulk * 4u, bar);
      }
    }
  }
  for (int i = nbulk + tid; i < span; i += NT) sx[i] = (i < nvalid) ? xc[(t0 + i) * p.x_st] : (TIn)0;
  for (int i = tid; i < nprog; i += NT) sprog[i] = p.words[w0 + i];
  __syncthreads();
  if (nbulk > 0) mbar_wait(bar, 0);

  const int warp = tid >> 5, lane = tid & 31;
  constexpr int kWarps = NT / 32;
  constexpr int kSub = 32 * R;
  const int nsub = p.tile / kSub;
  for (int sb = warp; sb < nsub; sb += kWarps) {
    const int base = sb * kSub;
    if (base >= remain) break;
    float yv[R];
    run_program<TIn, MODE, R, 32>(sx + base + lane, sprog, p.apply_gain, yv);
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const long long t = t0 + base + lane + 32 * r;
      if (t < p.frames) yc[t * p.y_st] = yv[r];
    }
  }
}

// Fallback for filters whose halo does not fit in shared memory: gather straight from global
// memory (L2 serves the reuse).  Same arithmetic, one output per thread-iteration.
template <typename TIn, int MODE>
__global__ void __launch_bounds__(256) fir_direct_kernel(const FirParams p) {
  const int c = blockIdx.y;
  const int w0 = p.offsets[c];
  const int nprog = p.offsets[c + 1] - w0;
  const TIn* __restrict__ xc = reinterpret_cast<const TIn*>(p.x) + (long long)c * p.x_sc;
  float* __restrict__ yc = p.y + (long long)c * p.y_sc;
  const int* __restrict__ prog = p.words + w0;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < p.frames; t += (long long)gridDim.x * blockDim.x) {
    if (nprog == 0) {
      yc[t * p.y_st] = (float)xc[t * p.x_st];
      continue;
    }
    float y = 0.0f;
    if constexpr (MODE == MODE_SEG) {
      const int S = prog[0];
      const int* seg = prog + 1;
      const int* tp = prog + 1 + 3 * S;
      for (int s = 0; s < S; ++s) {
        const int n_neg = seg[3 * s], n_pos = seg[3 * s + 1];
        float acc = 0.0f;
        for (int k = 0; k < n_neg; ++k) {
          const long long u = t + tp[k];
          acc = step_sub<TIn>(acc, u < p.frames ? xc[u * p.x_st] : (TIn)0);
        }
        tp += n_neg;
        for (int k = 0; k < n_pos; ++k) {
          const long long u = t + tp[k];
          acc = step_add<TIn>(acc, u < p.frames ? xc[u * p.x_st] : (TIn)0);
        }
        tp += n_pos;
        if (p.apply_gain) acc = fmul(acc, __int_as_float(seg[3 * s + 2]));
        y = fadd(y, acc);
      }
    } else {
      constexpr bool kF64 = is_f64<TIn>::value || MODE == MODE_ASC64;
      const int K = prog[0];
      const int* e = prog + 1;
      for (int k = 0; k < K; ++k) {
        const long long u = t + e[3 * k];
        const TIn xv = u < p.frames ? xc[u * p.x_st] : (TIn)0;
        const double coef = __hiloint2double(e[3 * k + 2], e[3 * k + 1]);
        if constexpr (kF64) y = __double2float_rn(dadd((double)y, dmul((double)xv, coef)));
        else y = fadd(y, fmul((float)xv, (float)coef));
      }
    }
    yc[t * p.y_st] = y;
  }
}

// ------------------------------------------------------------------------------------------------
// Fused stereo decorrelate: FIR on both channels of a frame, then M/S encode, width, gain and the
// Haas placement, one read of x and one write of out per sample.
// Shared memory: float x0[tile + halo] | float x1[tile + halo] | int prog0[] | int prog1[]
// ------------------------------------------------------------------------------------------------
struct StereoParams {
  const float* x;
  long long x_st, x_sc;
  void* out;
  long long o_st, o_sc;
  long long frames;
  const int* words;
  const int* offsets;  // 3 entries
  int tile, halo;
  int apply_gain;
  int ms_encode, use_width;
  float w_mid, w_side;  // float32(1 - width), float32(width)
  const float* gains;   // nullable, 2 floats
  int delay, delay_ch;
};

__device__ __forceinline__ void stereo_epilogue(float x0, float x1, float& y0, float& y1, const StereoParams& p, float g0, float g1) {
  if (p.ms_encode) {  // utils/dsp.py:59-63
    const float M = fadd(x0, x1);
    const float S = fmul(fsub(y0, y1), 0.5f);
    y0 = fmul(fadd(M, S), 0.5f);
    y1 = fmul(fsub(M, S), 0.5f);
  }
  if (p.use_width) {  // utils/dsp.py:34-37 with LR_to_MS / MS_to_LR (utils/dsp.py:140-166)
    float M = fmul(fadd(y0, y1), 0.5f);
    float S = fmul(fsub(y0, y1), 0.5f);
    M = fmul(M, p.w_mid);
    S = fmul(S, p.w_side);
    y0 = fadd(M, S);
    y1 = fsub(M, S);
  }
  if (p.gains) {
    y0 = fmul(y0, g0);
    y1 = fmul(y1, g1);
  }
}

template <typename TOut, int R>
__global__ void __launch_bounds__(256) vn_stereo_kernel(const StereoParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  constexpr int NT = 256;
  const int span = p.tile + p.halo;
  const int span_al = (span + 3) & ~3;
  float* s0 = reinterpret_cast<float*>(smem_raw);
  float* s1 = s0 + span_al;
  int* prog0 = reinterpret_cast<int*>(s1 + span_al);
  const int n0 = p.offsets[1] - p.offsets[0];
  const int n1 = p.offsets[2] - p.offsets[1];
  int* prog1 = prog0 + n0;

  const int tid = threadIdx.x;
  const long long t0 = (long long)blockIdx.x * p.tile;
  const long long remain = p.frames - t0;
  const int nvalid = (int)(remain < span ? remain : span);
  const float* __restrict__ x = p.x;
  for (int i = tid; i < span; i += NT) {
    float a = 0.0f, b = 0.0f;
    if (i < nvalid) {
      a = x[(t0 + i) * p.x_st];
      b = x[(t0 + i) * p.x_st + p.x_sc];
    }
    s0[i] = a;
    s1[i] = b;
  }
  for (int i = tid; i < n0 + n1; i += NT) prog0[i] = p.words[p.offsets[0] + i];
  __syncthreads();

  TOut* __restrict__ out = reinterpret_cast<TOut*>(p.out);
  const int sh0 = p.delay_ch == 0 ? p.delay : 0;
  const int sh1 = p.delay_ch == 1 ? p.delay : 0;
  float g0 = 1.0f, g1 = 1.0f;
  if (p.gains) {
    g0 = p.gains[0];
    g1 = p.gains[1];
  }
  const int warp = tid >> 5, lane = tid & 31;
  constexpr int kSub = 32 * R;
  const int nsub = p.tile / kSub;
  for (int sb = warp; sb < nsub; sb += NT / 32) {
    const int base = sb * kSub;
    if (base >= remain) break;
    float y0[R], y1[R];
    if (n0 > 0) run_program<float, MODE_SEG, R, 32>(s0 + base + lane, prog0, p.apply_gain, y0);
    if (n1 > 0) run_program<float, MODE_SEG, R, 32>(s1 + base + lane, prog1, p.apply_gain, y1);
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int i = base + lane + 32 * r;
      const long long t = t0 + i;
      if (t < p.frames) {
        const float a = s0[i], b = s1[i];
        float u = n0 > 0 ? y0[r] : a;  // unfiltered channel: copied through
        float v = n1 > 0 ? y1[r] : b;
        stereo_epilogue(a, b, u, v, p, g0, g1);
        out[(t + sh0) * p.o_st] = (TOut)u;
        out[(t + sh1) * p.o_st + p.o_sc] = (TOut)v;
      }
    }
  }
  // Haas zero fill: head of the delayed channel, tail of the other one (decorrelation.py:206-222)
  if (p.delay > 0) {
    if (blockIdx.x == 0) {
      for (int i = tid; i < p.delay; i += NT) out[(long long)i * p.o_st + (long long)p.delay_ch * p.o_sc] = (TOut)0;
    }
    if (blockIdx.x == gridDim.x - 1) {
      for (int i = tid; i < p.delay; i += NT) out[(p.frames + i) * p.o_st + (long long)(1 - p.delay_ch) * p.o_sc] = (TOut)0;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------------
constexpr int kR = 8;
// VND_DISABLE_WINDOW=1 in the environment forces the general tile kernel (A/B runs and tests).
static const bool g_disable_window = [] {
  const char* e = getenv("VND_DISABLE_WINDOW");
  return e && e[0] == '1';
}();

// VND_DISABLE_TMEM=1 skips the tensor-memory kernel (A/B runs and tests).
static const bool g_disable_tmem = [] {
  const char* e = getenv("VND_DISABLE_TMEM");
  return e && e[0] == '1';
}();

template <typename TIn, int MODE, int NT>
static int launch_tile(const FirParams& p, size_t smem, cudaStream_t st) {
  auto k = fir_tile_kernel<TIn, MODE, NT, kR>;
  VND_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long long blocks = (long long)p.tiles_per_channel * p.channels;
  VND_REQUIRE(blocks < 0x7fffffffLL, VND_EUNSUPPORTED, "grid too large (%lld CTAs)", blocks);
  k<<<(unsigned)blocks, NT, smem, st>>>(p);
  return after_launch("fir_tile_kernel");
}

template <typename TIn, int MODE>
static int launch_tile_nt(const FirParams& p, size_t smem, int nt, cudaStream_t st) {
  switch (nt) {
    case 256: return launch_tile<TIn, MODE, 256>(p, smem, st);
    case 512: return launch_tile<TIn, MODE, 512>(p, smem, st);
    default: return launch_tile<TIn, MODE, 1024>(p, smem, st);
  }
}

// Long filters (halo beyond 4096 samples, e.g. 300 impulses over 0.3 s @ 96 kHz: BASELINE config 4): tile + halo
// fills the SM's shared memory, so ONE CTA of 32 warps runs per SM.  R outputs per lane and pass; the tile is a whole
// number of passes for all 32 warps (round 1 ran 28 672 outputs = 3.5 passes of R = 8: half of the warps idled
// through the fourth pass), and a larger R spreads the tap-index load and the loop overhead of a tap over more outputs.
// Measured on 64 channels x 57.6 M frames (Gsamples/s, shared-memory pipe busy): round 1 (R = 8, 3.5 passes, scalar adds)
// 25.1 / 81 %; R = 8 25.3 / 81 %; R = 12 26.0 / 84 %; R = 16 (one pass of 16 384 outputs) 27.5 / 89 %.
// VND_LONG_R picks R in {8, 12, 16} (default 16).
static int long_filter_r() {
  static const int r = [] {
    const char* e = getenv("VND_LONG_R");
    const int v = e ? atoi(e) : 16;
    return (v == 8 || v == 12 || v == 16) ? v : 16;
  }();
  return r;
}
static int plan_long_tile(int halo, int max_prog_words, int r, size_t* smem) {
  const long long budget = (long long)kMaxDynSmem - 16 - 16 - (long long)max_prog_words * 4;
  const long long per_pass = 32LL * 32 * r;
  const long long tile = (budget / 4 - halo) / per_pass * per_pass;
  if (tile <= 0) return 0;
  *smem = 16 + (((size_t)(tile + halo) * 4 + 15) & ~(size_t)15) + (size_t)max_prog_words * 4;
  return (int)tile;
}
template <int R>
static int launch_long_tile(const FirParams& p, size_t smem, cudaStream_t st) {
  auto k = fir_tile_kernel<float, MODE_SEG, 1024, R>;
  VND_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long long blocks = (long long)p.tiles_per_channel * p.channels;
  VND_REQUIRE(blocks < 0x7fffffffLL, VND_EUNSUPPORTED, "grid too large (%lld CTAs)", blocks);
  k<<<(unsigned)blocks, 1024, smem, st>>>(p);
  return after_launch("fir_tile_kernel");
}

template <typename TIn, int MODE>
static int launch_direct(const FirParams& p, cudaStream_t st) {
  DeviceInfo di;
  int rc = device_info(&di);
  if (rc) return rc;
  long long bx = ceil_div<long long>(p.frames, 256);
  const long long cap = (long long)di.sm_count * 16;
  if (bx > cap) bx = cap;
  if (bx < 1) bx = 1;
  VND_REQUIRE(p.channels <= 65535, VND_EUNSUPPORTED, "direct FIR path supports at most 65535 channels");
  dim3 grid((unsigned)bx, (unsigned)p.channels);
  fir_direct_kernel<TIn, MODE><<<grid, 256, 0, st>>>(p);
  return after_launch("fir_direct_kernel");
}

// Picks the tile for a given halo so that tile + halo fits in shared memory; returns 0 when the
// halo itself does not fit (caller falls back to the direct kernel).
static int plan_tile(int halo, int elem, int max_prog_words, long long frames, int* nt, size_t* smem) {
  const long long budget = (long long)kMaxDynSmem - 16 - 16 - (long long)max_prog_words * 4;
  const long long span_max = budget / elem;
  constexpr int unit = 32 * kR;  // outputs per warp pass
  long long tile = 8192;
  if (halo > 4096) tile = ((long long)halo + 2047) / 2048 * 2048;  // keep the re-read of the halo <= 2x
  if (tile + halo > span_max) tile = (span_max - halo) / unit * unit;
  if (tile < unit) return 0;
  // no point in a tile longer than the signal
  const long long need = ceil_div<long long>(frames, unit) * unit;
  if (tile > need) tile = need;
  *smem = 16 + (((size_t)(tile + halo) * elem + 15) & ~(size_t)15) + (size_t)max_prog_words * 4;
  const int ctas_per_sm = (int)((228 * 1024) / (*smem + 1024));
  *nt = ctas_per_sm >= 4 ? 256 : (ctas_per_sm >= 2 ? 512 : 1024);
  // small tiles cannot feed many warps
  while (*nt > 256 && tile / unit < *nt / 32) *nt /= 2;
  return (int)tile;
}

int sparse_fir_launch(const vnd_signal* x, const vnd_signal* y, const vnd_tap_program* taps, int max_prog_words,
                      cudaStream_t st) {
  FirParams p{};
  p.x = x->data;
  p.x_st = x->stride_t;
  p.x_sc = x->stride_c;
  p.y = reinterpret_cast<float*>(y->data);
  p.y_st = y->stride_t;
  p.y_sc = y->stride_c;
  p.frames = x->frames;
  p.channels = taps->channels;
  p.words = taps->words;
  p.offsets = taps->offsets;
  p.apply_gain = taps->apply_gain;
  p.halo = taps->halo > 0 ? taps->halo : 0;
  if (p.frames == 0 || p.channels == 0) return VND_OK;
  const bool f64 = x->dtype == VND_F64;
  const int elem = f64 ? 8 : 4;
  int nt = 256;
  size_t smem = 0;
  // the halo never needs to reach past the end of the signal
  if (p.halo > p.frames) p.halo = (int)p.frames;
  p.halo = (p.halo + 3) & ~3;
  const int mode = taps->order == VND_ORDER_SEGMENTED ? MODE_SEG : (taps->order == VND_ORDER_ASCENDING ? MODE_ASC32 : MODE_ASC64);
  p.bulk_ok = (!f64 && x->stride_t == 1 && (x->stride_c % 4) == 0 && (reinterpret_cast<uintptr_t>(x->data) % 16) == 0) ? 1 : 0;
  if (!f64 && mode == MODE_SEG && !g_disable_tmem) {  // throughput path for long planar float32 slabs
    long long done = 0;
    const int rc = fir_tmem_launch(p, max_prog_words, st, &done);
    if (rc != VND_OK && rc != VND_EUNSUPPORTED) return rc;
    if (rc == VND_OK && done > 0) {  // the tail of every channel goes through the kernels below
      p.x = reinterpret_cast<const float*>(p.x) + done;
      p.y += done;
      p.frames -= done;
      if (p.frames == 0) return VND_OK;
      if (p.halo > p.frames) p.halo = (int)((p.frames + 3) & ~3LL);
      p.bulk_ok = (p.bulk_ok && (reinterpret_cast<uintptr_t>(p.x) % 16) == 0) ? 1 : 0;
    }
  }
  if (!f64 && mode == MODE_SEG && !g_disable_window) {  // planar float32 slabs too short for the above
    const int rc = fir_window_launch(p, max_prog_words, st);
    if (rc != VND_EUNSUPPORTED) return rc;
  }
  if (!f64 && mode == MODE_SEG && p.halo > 4096 && p.frames >= 4LL * p.halo) {  // long filters on long signals
    {  // the interior of every channel through the ring-buffer kernel (vnd_fir_ring.cu), the tail through the tile kernels below
      long long done = 0;
      const int rc = fir_ring_launch(p, max_prog_words, st, &done);
      if (rc != VND_OK && rc != VND_EUNSUPPORTED) return rc;
      if (rc == VND_OK && done > 0) {
        p.x = reinterpret_cast<const float*>(p.x) + done;
        p.y += done;
        p.frames -= done;
        if (p.frames == 0) return VND_OK;
        if (p.halo > p.frames) p.halo = (int)((p.frames + 3) & ~3LL);
        p.bulk_ok = (p.bulk_ok && (reinterpret_cast<uintptr_t>(p.x) % 16) == 0) ? 1 : 0;
      }
    }
  }
  if (!f64 && mode == MODE_SEG && p.halo > 4096 && p.frames >= 4LL * p.halo) {
    const int r = long_filter_r();
    const int tile = plan_long_tile(p.halo, max_prog_words, r, &smem);
    if (tile > 0) {
      p.tile = tile;
      p.tiles_per_channel = (int)ceil_div<long long>(p.frames, p.tile);
      if (r == 8) return launch_long_tile<8>(p, smem, st);
      if (r == 12) return launch_long_tile<12>(p, smem, st);
      return launch_long_tile<16>(p, smem, st);
    }
  }
  p.tile = plan_tile(p.halo, elem, max_prog_words, p.frames, &nt, &smem);
  if (p.tile == 0) {
    if (f64) {
      if (mode == MODE_SEG) return launch_direct<double, MODE_SEG>(p, st);
      return launch_direct<double, MODE_ASC64>(p, st);
    }
    if (mode == MODE_SEG) return launch_direct<float, MODE_SEG>(p, st);
    if (mode == MODE_ASC32) return launch_direct<float, MODE_ASC32>(p, st);
    return launch_direct<float, MODE_ASC64>(p, st);
  }
  p.tiles_per_channel = (int)ceil_div<long long>(p.frames, p.tile);
  if (f64) {
    if (mode == MODE_SEG) return launch_tile_nt<double, MODE_SEG>(p, smem, nt, st);
    return launch_tile_nt<double, MODE_ASC64>(p, smem, nt, st);
  }
  if (mode == MODE_SEG) return launch_tile_nt<float, MODE_SEG>(p, smem, nt, st);
  if (mode == MODE_ASC32) return launch_tile_nt<float, MODE_ASC32>(p, smem, nt, st);
  return launch_tile_nt<float, MODE_ASC64>(p, smem, nt, st);
}

// Returns VND_EUNSUPPORTED (without setting an error) when the fused kernel cannot hold the halo;
// the caller then composes the unfused kernels.
int vn_stereo_launch(const vnd_signal* x, void* out, int out_dtype, long long o_st, long long o_sc, const vnd_tap_program* taps,
                     int prog_words, const vnd_epilogue* ep, const float* gains, int delay, int delay_ch, cudaStream_t st) {
  StereoParams p{};
  p.x = reinterpret_cast<const float*>(x->data);
  p.x_st = x->stride_t;
  p.x_sc = x->stride_c;
  p.out = out;
  p.o_st = o_st;
  p.o_sc = o_sc;
  p.frames = x->frames;
  p.words = taps->words;
  p.offsets = taps->offsets;
  p.apply_gain = taps->apply_gain;
  p.ms_encode = ep->ms_encode;
  p.use_width = ep->use_width;
  p.w_mid = (float)(1.0 - (double)ep->width);
  p.w_side = (float)ep->width;
  p.gains = gains;
  p.delay = delay;
  p.delay_ch = delay_ch;
  int halo = taps->halo > 0 ? taps->halo : 0;
  if (halo > p.frames) halo = (int)p.frames;
  p.halo = (halo + 3) & ~3;
  constexpr int R = 4;
  constexpr int unit = 32 * R;
  long long tile = 1024;
  const long long need = ceil_div<long long>(p.frames, unit) * unit;
  if (tile > need) tile = need;
  size_t smem = (size_t)2 * (((size_t)tile + p.halo + 3) & ~(size_t)3) * 4 + (size_t)prog_words * 4 + 16;
  if (smem > (size_t)kMaxDynSmem) return VND_EUNSUPPORTED;
  p.tile = (int)tile;
  const unsigned blocks = (unsigned)ceil_div<long long>(p.frames, tile);
  if (out_dtype == VND_F64) {
    auto k = vn_stereo_kernel<double, R>;
    VND_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<blocks, 256, smem, st>>>(p);
  } else {
    auto k = vn_stereo_kernel<float, R>;
    VND_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<blocks, 256, smem, st>>>(p);
  }
  return after_launch("vn_stereo_kernel");
}

}  // namespace vnd
