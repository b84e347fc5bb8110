// Batched stereo-image objective kernels (sm_100a).
//
// Reference arithmetic evaluated here (paths relative to the reference root):
//   symmetry_aware_objective   src/vndecorrelate/optimization.py:46-105
//   polar_coordinates          src/vndecorrelate/utils/dsp.py:374-422  (mode 'MS', semicircular)
//   moment helpers             src/vndecorrelate/optimization.py:11-43
//   candidates                 src/vndecorrelate/optimization.py:260-272 (velvet noise, channel 0
//                              filtered, channel 1 copied, LR, no normaliser) and :183-203 (Haas)
//
// The kernels produce per (clip, candidate) the sums the objective is made of; the host applies the
// reference's scalar dtype chain.  Two facts shape the velvet-noise kernel:
//   * folded theta == atan(d / s) with d = L - R, s = L + R, so no atan2 + fold is needed for the
//     amplitude-weighted sums, whose weights make an absolute error of 1e-7 rad irrelevant;
//   * max|theta| IS ulp-sensitive (the penalty term multiplies it by ~1.6e3), but rounding is
//     monotone, so max|theta| is attained at the frame with the largest |d|/|s| among frames with
//     s >= 0 and among frames with s < 0.  The kernel tracks those two frames exactly and the host
//     evaluates a correctly rounded float32 arctan2 for just those.
//
// Decomposition: CTA = (chunk of tiles, clip, candidate group).  A tile of both channels (+ the
// filter halo for channel 0) is staged in shared memory once and reused by every candidate of the
// group; each WARP owns whole candidates (no block-level reduction, no __syncthreads in the
// candidate loop), reduces with warp shuffles and accumulates float64 partials in shared memory.

#include "vnd_common.cuh"

namespace vnd {

constexpr int OBJ_SLOTS = 12;  // doubles per (clip, candidate) partial, see vnd_b200.h
constexpr int OBJ_NT = 512;
constexpr int OBJ_R = 8;
constexpr int OBJ_TILE = 4096;

struct ObjParams {
  const float* clips;
  long long frames, clip_stride, chan_stride;
  const int* words;
  const int* offsets;
  int n_cand;
  int apply_gain;
  int halo;
  int cand_per_group;
  int tiles_per_chunk;
  int n_chunks;
  double* chunk_partials;  // [clip][chunk][cand][OBJ_SLOTS]
};

// atan(t) for t in [0, 1]: t * P(t^2), |error| <= 1e-7 (degree-8 minimax fit, float32 Horner).
__device__ __forceinline__ float atan01(float t) {
  const float z = t * t;
  float p = 0.00245671847107214f;
  p = fmaf(p, z, -0.01440133168500584f);
  p = fmaf(p, z, 0.03978117728736144f);
  p = fmaf(p, z, -0.07234853052703884f);
  p = fmaf(p, z, 0.10498943808759016f);
  p = fmaf(p, z, -0.14161228535203682f);
  p = fmaf(p, z, 0.19985906672823953f);
  p = fmaf(p, z, -0.3333259702410447f);
  p = fmaf(p, z, 0.9999998863844667f);
  return p * t;
}

struct LaneAcc {
  float sr, srt, srt2, srt3, slr, sll;
  float d_pos, s_pos, d_neg, s_neg;  // |d|, |s| of the frame with the largest |d|/|s| per sign of s
};

__device__ __forceinline__ void lane_acc_frame(LaneAcc& a, float l, float r_) {
  const float d = fsub(l, r_), s = fadd(l, r_);  // utils/dsp.py:399-401
  const float ad = fabsf(d), as = fabsf(s);
  const float mx = fmaxf(ad, as), mn = fminf(ad, as);
  const float t = mx > 0.0f ? __fdividef(mn, mx) : 0.0f;
  float th = atan01(t);
  if (ad > as) th = 1.57079632679489662f - th;
  th = __int_as_float(__float_as_int(th) | ((__float_as_int(d) ^ __float_as_int(s)) & 0x80000000));
  const float rad = __fsqrt_rn(fadd(fmul(l, l), fmul(r_, r_)));  // utils/dsp.py:413
  const float rt = rad * th;
  a.sr += rad;
  a.srt += rt;
  a.srt2 = fmaf(rt, th, a.srt2);
  a.srt3 = fmaf(rt * th, th, a.srt3);
  a.slr = fmaf(l, r_, a.slr);
  a.sll = fmaf(l, l, a.sll);
  // exact-enough ordering of the ratios by cross multiplication (ties keep the earlier frame)
  if (s >= 0.0f) {
    if (ad * a.s_pos > a.d_pos * as) { a.d_pos = ad; a.s_pos = as; }
  } else {
    if (ad * a.s_neg > a.d_neg * as) { a.d_neg = ad; a.s_neg = as; }
  }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// best ratio across the warp: every lane ends with the same (d, s)
__device__ __forceinline__ void warp_best_ratio(float& d, float& s) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float od = __shfl_xor_sync(0xffffffffu, d, o);
    const float os = __shfl_xor_sync(0xffffffffu, s, o);
    const double lhs = (double)od * (double)s, rhs = (double)d * (double)os;  // exact products
    if (lhs > rhs || (lhs == rhs && od > d)) { d = od; s = os; }
  }
}

template <int R>
__device__ __forceinline__ void run_candidate(const float* __restrict__ px, const int* __restrict__ prog, int apply_gain, float (&yv)[R]) {
#pragma unroll
  for (int r = 0; r < R; ++r) yv[r] = 0.0f;
  const int S = prog[0];
  const int* seg = prog + 1;
  const int* tp = prog + 1 + 3 * S;
  for (int s = 0; s < S; ++s) {
    const int n_neg = seg[3 * s], n_pos = seg[3 * s + 1];
    const float gain = __int_as_float(seg[3 * s + 2]);
    float acc[R];
#pragma unroll
    for (int r = 0; r < R; ++r) acc[r] = 0.0f;
    for (int k = 0; k < n_neg; ++k) {
      const float* q = px + tp[k];
#pragma unroll
      for (int r = 0; r < R; ++r) acc[r] = fsub(acc[r], q[32 * r]);
    }
    tp += n_neg;
    for (int k = 0; k < n_pos; ++k) {
      const float* q = px + tp[k];
#pragma unroll
      for (int r = 0; r < R; ++r) acc[r] = fadd(acc[r], q[32 * r]);
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

// Shared memory: float x0[TILE + halo] | float x1[TILE] | double acc[cand_per_group][OBJ_SLOTS]
//                | int prog[warps][max_prog_words]
__global__ void __launch_bounds__(OBJ_NT) vn_objective_kernel(const ObjParams p, int max_prog_words) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int span = OBJ_TILE + p.halo;
  float* s0 = reinterpret_cast<float*>(smem_raw);
  float* s1 = s0 + span;
  double* acc = reinterpret_cast<double*>(s1 + OBJ_TILE);
  int* progs = reinterpret_cast<int*>(acc + (size_t)p.cand_per_group * OBJ_SLOTS);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr int kWarps = OBJ_NT / 32;
  const int chunk = blockIdx.x, clip = blockIdx.y, group = blockIdx.z;
  const int cand0 = group * p.cand_per_group;
  const int ncand = min(p.cand_per_group, p.n_cand - cand0);
  const float* __restrict__ x0 = p.clips + (long long)clip * p.clip_stride;
  const float* __restrict__ x1 = x0 + p.chan_stride;
  int* myprog = progs + warp * max_prog_words;

  for (int i = tid; i < ncand * OBJ_SLOTS; i += OBJ_NT) {
    const int slot = i % OBJ_SLOTS;
    acc[i] = (slot == 7 || slot == 9) ? 1.0 : 0.0;  // ratio trackers start at 0 / 1
  }

  const long long tile_first = (long long)chunk * p.tiles_per_chunk;
  for (int ti = 0; ti < p.tiles_per_chunk; ++ti) {
    const long long t0 = (tile_first + ti) * OBJ_TILE;
    if (t0 >= p.frames) break;
    const long long remain = p.frames - t0;
    __syncthreads();  // previous tile fully consumed (also orders the acc init)
    for (int i = tid; i < span; i += OBJ_NT) s0[i] = i < remain ? x0[t0 + i] : 0.0f;
    for (int i = tid; i < OBJ_TILE; i += OBJ_NT) s1[i] = i < remain ? x1[t0 + i] : 0.0f;
    __syncthreads();
    const int nvalid = (int)(remain < OBJ_TILE ? remain : OBJ_TILE);

    for (int ci = warp; ci < ncand; ci += kWarps) {
      const int w0 = p.offsets[cand0 + ci];
      const int nprog = p.offsets[cand0 + ci + 1] - w0;
      __syncwarp();
      for (int i = lane; i < nprog; i += 32) myprog[i] = p.words[w0 + i];
      __syncwarp();
      LaneAcc a{0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 1.f, 0.f, 1.f};
      for (int base = 0; base < nvalid; base += 32 * OBJ_R) {
        float yv[OBJ_R];
        run_candidate<OBJ_R>(s0 + base + lane, myprog, p.apply_gain, yv);
#pragma unroll
        for (int r = 0; r < OBJ_R; ++r) {
          const int i = base + lane + 32 * r;
          if (i < nvalid) lane_acc_frame(a, yv[r], s1[i]);
        }
      }
      a.sr = warp_sum(a.sr);
      a.srt = warp_sum(a.srt);
      a.srt2 = warp_sum(a.srt2);
      a.srt3 = warp_sum(a.srt3);
      a.slr = warp_sum(a.slr);
      a.sll = warp_sum(a.sll);
      warp_best_ratio(a.d_pos, a.s_pos);
      warp_best_ratio(a.d_neg, a.s_neg);
      if (lane == 0) {
        double* q = acc + (size_t)ci * OBJ_SLOTS;
        q[0] += (double)a.sr;
        q[1] += (double)a.srt;
        q[2] += (double)a.srt2;
        q[3] += (double)a.srt3;
        q[4] += (double)a.slr;
        q[5] += (double)a.sll;
        if ((double)a.d_pos * q[7] > q[6] * (double)a.s_pos) { q[6] = a.d_pos; q[7] = a.s_pos; }
        if ((double)a.d_neg * q[9] > q[8] * (double)a.s_neg) { q[8] = a.d_neg; q[9] = a.s_neg; }
        q[10] += (double)nvalid;
      }
    }
  }
  __syncthreads();
  double* out = p.chunk_partials + (((size_t)clip * p.n_chunks + chunk) * p.n_cand + cand0) * OBJ_SLOTS;
  for (int i = tid; i < ncand * OBJ_SLOTS; i += OBJ_NT) out[i] = acc[i];
}

// Fixed-order combine of the chunk partials: sums added chunk by chunk, ratio trackers by exact
// cross multiplication (earlier chunk wins ties, like np.max keeps the value either way).
__global__ void obj_combine_kernel(const double* __restrict__ chunk_partials, double* __restrict__ partials, int n_clips,
                                   int n_chunks, int n_cand) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)n_clips * n_cand) return;
  const int clip = (int)(idx / n_cand), cand = (int)(idx % n_cand);
  double r[OBJ_SLOTS];
  for (int k = 0; k < OBJ_SLOTS; ++k) r[k] = 0.0;
  r[7] = 1.0;
  r[9] = 1.0;
  for (int ch = 0; ch < n_chunks; ++ch) {
    const double* q = chunk_partials + (((size_t)clip * n_chunks + ch) * n_cand + cand) * OBJ_SLOTS;
    for (int k = 0; k < 6; ++k) r[k] += q[k];
    if (q[6] * r[7] > r[6] * q[7]) { r[6] = q[6]; r[7] = q[7]; }
    if (q[8] * r[9] > r[8] * q[9]) { r[8] = q[8]; r[9] = q[9]; }
    r[10] += q[10];
  }
  double* o = partials + (size_t)idx * OBJ_SLOTS;
  for (int k = 0; k < OBJ_SLOTS; ++k) o[k] = r[k];
}

// ------------------------------------------------------------------------------------------------
// Haas candidates: float64 throughout (optimization.py:183-203 builds LR-mode HaasEffects whose
// output is float64).  One CTA per (clip, candidate); 8 doubles per pair:
//   [sum r, sum r*th, sum r*th^2, sum r*th^3, max|th|, sum L*R, sum L*L, frames]
// ------------------------------------------------------------------------------------------------
constexpr int HAAS_SLOTS = 8;

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <typename TIn>
__global__ void __launch_bounds__(256) haas_objective_kernel(const TIn* __restrict__ clips, long long frames, long long clip_stride,
                                                             long long chan_stride, const int* __restrict__ delays, int n_cand,
                                                             double* __restrict__ partials) {
  const int cand = blockIdx.x, clip = blockIdx.y;
  const int d = delays[cand];
  const TIn* __restrict__ x0 = clips + (long long)clip * clip_stride;
  const TIn* __restrict__ x1 = x0 + chan_stride;
  const long long total = frames + d;
  const double kHalfPi = 1.5707963267948966, kPi = 3.141592653589793;
  double v[7] = {0, 0, 0, 0, 0, 0, 0};
  for (long long m = threadIdx.x; m < total; m += blockDim.x) {
    const double l = (m >= d) ? (double)x0[m - d] : 0.0;  // channel 0 delayed (decorrelation.py:220-222)
    const double r = (m < frames) ? (double)x1[m] : 0.0;
    double th = atan2(l - r, l + r);
    if (th < -kHalfPi) th += kPi;
    else if (th > kHalfPi) th -= kPi;
    const double rad = sqrt(l * l + r * r);
    v[0] += rad;
    v[1] += rad * th;
    v[2] += rad * (th * th);
    v[3] += rad * (th * th * th);
    v[4] = fmax(v[4], fabs(th));
    v[5] += l * r;
    v[6] += l * l;
  }
  __shared__ double red[8][7];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int k = 0; k < 7; ++k) {
    if (k == 4) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v[4] = fmax(v[4], __shfl_xor_sync(0xffffffffu, v[4], o));
    } else {
      v[k] = warp_sum_d(v[k]);
    }
  }
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < 7; ++k) red[warp][k] = v[k];
  }
  __syncthreads();
  if (threadIdx.x < 7) {
    const int k = threadIdx.x;
    double r = red[0][k];
    for (int w = 1; w < 8; ++w) r = (k == 4) ? fmax(r, red[w][k]) : r + red[w][k];
    partials[((size_t)clip * n_cand + cand) * HAAS_SLOTS + k] = r;
  }
  if (threadIdx.x == 7) partials[((size_t)clip * n_cand + cand) * HAAS_SLOTS + 7] = (double)total;
}

// ------------------------------------------------------------------------------------------------
// launch planning
// ------------------------------------------------------------------------------------------------
struct ObjPlan {
  int cand_per_group, n_groups, tiles_per_chunk, n_chunks;
  size_t smem;
};

static int plan_objective(long long frames, int n_clips, int n_cand, int halo, int max_prog_words, ObjPlan* pl) {
  DeviceInfo di;
  int rc = device_info(&di);
  if (rc) return rc;
  const size_t fixed = (size_t)(OBJ_TILE + halo + OBJ_TILE) * 4 + (size_t)(OBJ_NT / 32) * max_prog_words * 4 + 64;
  if (fixed + OBJ_SLOTS * 8 * 16 > (size_t)kMaxDynSmem) return VND_EUNSUPPORTED;
  int cpg = (int)(((size_t)kMaxDynSmem - fixed) / (OBJ_SLOTS * 8));
  if (cpg > n_cand) cpg = n_cand;
  if (cpg > 1024) cpg = 1024;
  const long long tiles = ceil_div<long long>(frames, OBJ_TILE);
  // enough CTAs for ~4 waves of one CTA per SM: first split tiles into chunks, then candidates
  const long long want = (long long)di.sm_count * 4;
  long long n_chunks = ceil_div<long long>(want, (long long)n_clips * ceil_div(n_cand, cpg));
  if (n_chunks > tiles) n_chunks = tiles;
  if (n_chunks < 1) n_chunks = 1;
  int tpc = (int)ceil_div<long long>(tiles, n_chunks);
  n_chunks = ceil_div<long long>(tiles, tpc);
  while ((long long)n_clips * n_chunks * ceil_div(n_cand, cpg) < want && cpg > 16) cpg = (cpg + 1) / 2;
  pl->cand_per_group = cpg;
  pl->n_groups = ceil_div(n_cand, cpg);
  pl->tiles_per_chunk = tpc;
  pl->n_chunks = (int)n_chunks;
  pl->smem = fixed + (size_t)cpg * OBJ_SLOTS * 8;
  return VND_OK;
}

int objective_workspace_bytes(long long frames, int n_clips, int n_cand, size_t* bytes) {
  // upper bound independent of the program: one partial block per (clip, tile, candidate) is never
  // exceeded because n_chunks <= tiles; cap chunks at what plan_objective can ask for.
  DeviceInfo di;
  int rc = device_info(&di);
  if (rc) return rc;
  const long long tiles = ceil_div<long long>(frames, OBJ_TILE);
  long long chunks = (long long)di.sm_count * 4;
  if (chunks > tiles) chunks = tiles;
  if (chunks < 1) chunks = 1;
  *bytes = (size_t)n_clips * chunks * n_cand * OBJ_SLOTS * 8 + 256;
  return VND_OK;
}

int vn_objective_launch(const float* clips, long long frames, int n_clips, long long clip_stride, long long chan_stride,
                        const vnd_tap_program* cand, double* partials, void* workspace, size_t workspace_bytes, cudaStream_t st) {
  if (n_clips == 0 || cand->channels == 0) return VND_OK;
  int halo = cand->halo > 0 ? cand->halo : 0;
  if (halo > frames) halo = (int)frames;
  halo = (halo + 3) & ~3;
  const int mpw = cand->max_channel_words > 0 ? cand->max_channel_words : 1;
  ObjPlan pl;
  int rc = plan_objective(frames, n_clips, cand->channels, halo, mpw, &pl);
  if (rc == VND_EUNSUPPORTED) set_error("objective kernel: filter halo %d does not fit in shared memory", halo);
  if (rc) return rc;
  const size_t need = (size_t)n_clips * pl.n_chunks * cand->channels * OBJ_SLOTS * 8;
  VND_REQUIRE(workspace && workspace_bytes >= need, VND_ENOMEM, "objective workspace too small: need %zu bytes, have %zu", need,
              workspace_bytes);
  VND_REQUIRE(n_clips <= 65535 && pl.n_groups <= 65535, VND_EUNSUPPORTED, "too many clips/candidate groups for one launch");
  ObjParams p{};
  p.clips = clips;
  p.frames = frames;
  p.clip_stride = clip_stride;
  p.chan_stride = chan_stride;
  p.words = cand->words;
  p.offsets = cand->offsets;
  p.n_cand = cand->channels;
  p.apply_gain = cand->apply_gain;
  p.halo = halo;
  p.cand_per_group = pl.cand_per_group;
  p.tiles_per_chunk = pl.tiles_per_chunk;
  p.n_chunks = pl.n_chunks;
  p.chunk_partials = reinterpret_cast<double*>(workspace);
  VND_CUDA_OK(cudaFuncSetAttribute(vn_objective_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
  dim3 grid((unsigned)pl.n_chunks, (unsigned)n_clips, (unsigned)pl.n_groups);
  vn_objective_kernel<<<grid, OBJ_NT, pl.smem, st>>>(p, mpw);
  rc = after_launch("vn_objective_kernel");
  if (rc) return rc;
  const long long pairs = (long long)n_clips * cand->channels;
  obj_combine_kernel<<<(unsigned)ceil_div<long long>(pairs, 128), 128, 0, st>>>(p.chunk_partials, partials, n_clips, pl.n_chunks,
                                                                                 cand->channels);
  return after_launch("obj_combine_kernel");
}

int haas_objective_launch(const void* clips, int clip_dtype, long long frames, int n_clips, long long clip_stride, long long chan_stride,
                          const int* delays, int n_cand, double* partials, cudaStream_t st) {
  if (n_clips == 0 || n_cand == 0) return VND_OK;
  VND_REQUIRE(n_clips <= 65535, VND_EUNSUPPORTED, "too many clips for one launch");
  dim3 grid((unsigned)n_cand, (unsigned)n_clips);
  if (clip_dtype == VND_F64)
    haas_objective_kernel<double><<<grid, 256, 0, st>>>((const double*)clips, frames, clip_stride, chan_stride, delays, n_cand, partials);
  else
    haas_objective_kernel<float><<<grid, 256, 0, st>>>((const float*)clips, frames, clip_stride, chan_stride, delays, n_cand, partials);
  return after_launch("haas_objective_kernel");
}

}  // namespace vnd
