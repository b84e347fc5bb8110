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
//     s >= 0 and among frames with s < 0.  The kernel tracks those two frames exactly (track_ratio: the cross
//     products are compared in float64, where they are exact) and the host evaluates a correctly rounded float32
//     arctan2 for just those.
//
// Decomposition: CTA = (chunk of tiles, clip, candidate group).  A tile of both channels (+ the
// filter halo for channel 0) is staged in shared memory once and reused by every candidate of the
// group; each WARP owns whole candidates (no block-level reduction, no __syncthreads in the
// candidate loop), reduces with warp shuffles and accumulates float64 partials in shared memory.

#include <stdlib.h>

#include "vnd_objective.cuh"

namespace vnd {

size_t objective_tmem_workspace_bytes(long long frames, int n_clips, int n_cand, int sm_count);
int vn_objective_tmem_launch(const float* clips, long long frames, int n_clips, long long clip_stride, long long chan_stride,
                             const vnd_tap_program* cand, double* partials, void* workspace, size_t workspace_bytes, cudaStream_t st);

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
      for (int r = 0; r < R; r += 2) obj_sub2(acc[r], acc[r + 1], q[32 * r], q[32 * r + 32]);
    }
    tp += n_neg;
    for (int k = 0; k < n_pos; ++k) {
      const float* q = px + tp[k];
#pragma unroll
      for (int r = 0; r < R; r += 2) obj_add2(acc[r], acc[r + 1], q[32 * r], q[32 * r + 32]);
    }
    tp += n_pos;
    if (apply_gain) {
#pragma unroll
      for (int r = 0; r < R; r += 2) obj_mul2(acc[r], acc[r + 1], gain, gain);
    }
#pragma unroll
    for (int r = 0; r < R; r += 2) obj_add2(yv[r], yv[r + 1], acc[r], acc[r + 1]);
  }
}

// Whole tiles: a lane owns OBJ_R / 4 groups of FOUR CONSECUTIVE frames (frames base + 128 j + 4 lane + e),
// and the filtered channel is staged four times, copy c shifted by c samples, so that every tap finds a
// copy in which its operands are 16-byte aligned: one LDS.128 per group and tap instead of four LDS.32
// (same shared-memory wavefronts, a quarter of the load instructions).  `dec` holds the program's taps
// decoded once per candidate as word offsets (offset & 3) * span4 + (offset & ~3) into the copies.
#ifndef VND_OBJ_DUAL
#define VND_OBJ_DUAL 0  // 1: two decay segments interleaved per warp (needs 128 registers: VND_OBJ_NT <= 512).  Measured on 16 clips x 1024
                        // strengths x 30 s: 129 k evaluations/s with 16 warps against 156 k for the single stream with 20 warps (147 k
                        // with 16): unlike in fir_ring_kernel, the warps the second stream costs are worth more to the polar moments.
#endif
template <int R, bool SUB>
__device__ __forceinline__ void obj_tap_v4(const float* __restrict__ px, int off, float (&acc)[R]) {
  const float4* q = reinterpret_cast<const float4*>(px + off);
#pragma unroll
  for (int j = 0; j < R / 4; ++j) {
    const float4 v = q[32 * j];
    if constexpr (SUB) {
      obj_sub2(acc[4 * j], acc[4 * j + 1], v.x, v.y);
      obj_sub2(acc[4 * j + 2], acc[4 * j + 3], v.z, v.w);
    } else {
      obj_add2(acc[4 * j], acc[4 * j + 1], v.x, v.y);
      obj_add2(acc[4 * j + 2], acc[4 * j + 3], v.z, v.w);
    }
  }
}
template <int R>
__device__ __forceinline__ void run_candidate_v4(const float* __restrict__ px, const int* __restrict__ prog, const int* __restrict__ dec,
                                                 int apply_gain, float (&yv)[R]) {
#pragma unroll
  for (int r = 0; r < R; ++r) yv[r] = 0.0f;
  const int S = prog[0];
  const int* seg = prog + 1;
  const int* tp = dec;
#if VND_OBJ_DUAL
  // Two decay segments at a time (their sums are independent, optimization.py's candidates go through decorrelation.py:402-414):
  // the loads of both taps are issued before either's adds, which doubles the loads a warp has in flight; the running output
  // still adds the scaled sums in segment order.  Offsets are read one tap ahead (a word of slack follows the decoded list).
  for (int s = 0; s < S; s += 2) {
    const bool two = s + 1 < S;
    const int na_neg = seg[3 * s], na = na_neg + seg[3 * s + 1];
    const int nb_neg = two ? seg[3 * s + 3] : 0, nb = two ? nb_neg + seg[3 * s + 4] : 0;
    const int* ta = tp;
    const int* tb = tp + na;
    tp += na + nb;
    float accA[R], accB[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      accA[r] = 0.0f;
      accB[r] = 0.0f;
    }
    const int nboth = na < nb ? na : nb;
    int oa = ta[0], ob = tb[0];
    int k = 0;
    for (; k < nboth; ++k) {
      const float4* qa = reinterpret_cast<const float4*>(px + oa);
      const float4* qb = reinterpret_cast<const float4*>(px + ob);
      oa = ta[k + 1];
      ob = tb[k + 1];
      float4 va[R / 4], vb[R / 4];
#pragma unroll
      for (int j = 0; j < R / 4; ++j) va[j] = qa[32 * j];
#pragma unroll
      for (int j = 0; j < R / 4; ++j) vb[j] = qb[32 * j];
      if (k < na_neg) {
#pragma unroll
        for (int j = 0; j < R / 4; ++j) {
          obj_sub2(accA[4 * j], accA[4 * j + 1], va[j].x, va[j].y);
          obj_sub2(accA[4 * j + 2], accA[4 * j + 3], va[j].z, va[j].w);
        }
      } else {
#pragma unroll
        for (int j = 0; j < R / 4; ++j) {
          obj_add2(accA[4 * j], accA[4 * j + 1], va[j].x, va[j].y);
          obj_add2(accA[4 * j + 2], accA[4 * j + 3], va[j].z, va[j].w);
        }
      }
      if (k < nb_neg) {
#pragma unroll
        for (int j = 0; j < R / 4; ++j) {
          obj_sub2(accB[4 * j], accB[4 * j + 1], vb[j].x, vb[j].y);
          obj_sub2(accB[4 * j + 2], accB[4 * j + 3], vb[j].z, vb[j].w);
        }
      } else {
#pragma unroll
        for (int j = 0; j < R / 4; ++j) {
          obj_add2(accB[4 * j], accB[4 * j + 1], vb[j].x, vb[j].y);
          obj_add2(accB[4 * j + 2], accB[4 * j + 3], vb[j].z, vb[j].w);
        }
      }
    }
    for (int ka = k; ka < na; ++ka) {
      if (ka < na_neg) obj_tap_v4<R, true>(px, ta[ka], accA);
      else obj_tap_v4<R, false>(px, ta[ka], accA);
    }
    for (int kb = k; kb < nb; ++kb) {
      if (kb < nb_neg) obj_tap_v4<R, true>(px, tb[kb], accB);
      else obj_tap_v4<R, false>(px, tb[kb], accB);
    }
    if (apply_gain) {
      const float ga = __int_as_float(seg[3 * s + 2]);
#pragma unroll
      for (int r = 0; r < R; r += 2) obj_mul2(accA[r], accA[r + 1], ga, ga);
    }
#pragma unroll
    for (int r = 0; r < R; r += 2) obj_add2(yv[r], yv[r + 1], accA[r], accA[r + 1]);
    if (two) {
      if (apply_gain) {
        const float gb = __int_as_float(seg[3 * s + 5]);
#pragma unroll
        for (int r = 0; r < R; r += 2) obj_mul2(accB[r], accB[r + 1], gb, gb);
      }
#pragma unroll
      for (int r = 0; r < R; r += 2) obj_add2(yv[r], yv[r + 1], accB[r], accB[r + 1]);
    }
  }
#else
  for (int s = 0; s < S; ++s) {
    const int n_neg = seg[3 * s], n_pos = seg[3 * s + 1];
    const float gain = __int_as_float(seg[3 * s + 2]);
    float acc[R];
#pragma unroll
    for (int r = 0; r < R; ++r) acc[r] = 0.0f;
    int nxt = tp[0];  // offsets are read one tap ahead (a word of slack follows the decoded list)
    for (int k = 0; k < n_neg; ++k) {
      const int off = nxt;
      nxt = tp[k + 1];
      obj_tap_v4<R, true>(px, off, acc);
    }
    tp += n_neg;
    for (int k = 0; k < n_pos; ++k) {
      const int off = nxt;
      nxt = tp[k + 1];
      obj_tap_v4<R, false>(px, off, acc);
    }
    tp += n_pos;
    if (apply_gain) {
#pragma unroll
      for (int r = 0; r < R; r += 2) obj_mul2(acc[r], acc[r + 1], gain, gain);
    }
#pragma unroll
    for (int r = 0; r < R; r += 2) obj_add2(yv[r], yv[r + 1], acc[r], acc[r + 1]);
  }
#endif
}

// Shared memory: float x0[4][span4] (copy c = channel 0 shifted by c samples) | float x1[TILE] |
//                double acc[cand_per_group][OBJ_SLOTS] | int prog[warps][max_prog_words] | int dec[warps][max_prog_words]
#ifndef VND_OBJ_MINB
#define VND_OBJ_MINB 1  // CTAs per SM the register and shared-memory budgets are planned for (measured: 16 frames per lane in one CTA beats 8 in two)
#endif
__global__ void __launch_bounds__(OBJ_NT, VND_OBJ_MINB) vn_objective_kernel(const ObjParams p, int max_prog_words) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int span = OBJ_TILE + p.halo;
  const int span4 = (span + 7) & ~3;  // words per shifted copy (a multiple of 4, room for the shift)
  float* s0 = reinterpret_cast<float*>(smem_raw);
  float* s1 = s0 + 4 * span4;
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
  int* mydec = progs + (kWarps + warp) * max_prog_words;

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
    for (int i = tid; i < span + 3; i += OBJ_NT) {  // copy c holds x0[t0 + c + j] at j
      const float v = i < remain ? x0[t0 + i] : 0.0f;
      if (i < span) s0[i] = v;
      if (i >= 1 && i - 1 < span) s0[span4 + i - 1] = v;
      if (i >= 2 && i - 2 < span) s0[2 * span4 + i - 2] = v;
      if (i >= 3) s0[3 * span4 + i - 3] = v;
    }
    for (int i = tid; i < OBJ_TILE; i += OBJ_NT) s1[i] = i < remain ? x1[t0 + i] : 0.0f;
    __syncthreads();
    const int nvalid = (int)(remain < OBJ_TILE ? remain : OBJ_TILE);

    for (int ci = warp; ci < ncand; ci += kWarps) {
      const int w0 = p.offsets[cand0 + ci];
      const int nprog = p.offsets[cand0 + ci + 1] - w0;
      __syncwarp();
      for (int i = lane; i < nprog; i += 32) myprog[i] = p.words[w0 + i];
      __syncwarp();
      {  // taps as word offsets into the shifted copies
        const int ntap0 = 1 + 3 * myprog[0];
        for (int i = ntap0 + lane; i < nprog; i += 32) {
          const int off = myprog[i];
          mydec[i - ntap0] = (off & 3) * span4 + (off & ~3);
        }
      }
      __syncwarp();
      LaneAcc a{0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 1.f, 0.f, 1.f};
      if (nvalid == OBJ_TILE) {  // whole tile: no bounds checks, frames in packed pairs
        LaneAcc2 b{{0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}, 0.f, 1.f, 0.f, 1.f, -1.f, -1.f};
        for (int base = 0; base < OBJ_TILE; base += 32 * OBJ_R) {
          float yv[OBJ_R];
          run_candidate_v4<OBJ_R>(s0 + base + 4 * lane, myprog, mydec, p.apply_gain, yv);
          const float4* r1 = reinterpret_cast<const float4*>(s1 + base + 4 * lane);
#pragma unroll
          for (int j = 0; j < OBJ_R / 4; ++j) {
            const float4 v = r1[32 * j];
            lane_acc_pair(b, yv[4 * j], yv[4 * j + 1], v.x, v.y);
            lane_acc_pair(b, yv[4 * j + 2], yv[4 * j + 3], v.z, v.w);
          }
        }
        a.sr = b.sr[0] + b.sr[1];
        a.srt = b.srt[0] + b.srt[1];
        a.srt2 = b.srt2[0] + b.srt2[1];
        a.srt3 = b.srt3[0] + b.srt3[1];
        a.slr = b.slr[0] + b.slr[1];
        a.sll = b.sll[0] + b.sll[1];
        a.d_pos = b.d_pos;
        a.s_pos = b.s_pos;
        a.d_neg = b.d_neg;
        a.s_neg = b.s_neg;
      } else {
        for (int base = 0; base < nvalid; base += 32 * OBJ_R) {
          float yv[OBJ_R];
          run_candidate<OBJ_R>(s0 + base + lane, myprog, p.apply_gain, yv);
#pragma unroll
          for (int r = 0; r < OBJ_R; ++r) {
            const int i = base + lane + 32 * r;
            if (i < nvalid) lane_acc_frame(a, yv[r], s1[i]);
          }
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

int obj_combine_launch(const double* chunk_partials, double* partials, int n_clips, int n_chunks, int n_cand, cudaStream_t st) {
  const long long pairs = (long long)n_clips * n_cand;
  obj_combine_kernel<<<(unsigned)ceil_div<long long>(pairs, 128), 128, 0, st>>>(chunk_partials, partials, n_clips, n_chunks, n_cand);
  return after_launch("obj_combine_kernel");
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

// The float64 pipe is what bounds this kernel (one atan2, one sqrt and the moments per frame and delay; the clip itself
// stays in L2 across its delays), so both functions are written out lean instead of calling the library versions with
// their special-case paths: ~45 instead of ~100 float64 instructions per frame.
//
// Folded angle (utils/dsp.py:399-412): atan2(d, s) brought into [-pi/2, pi/2] by adding or subtracting pi is atan(d / s)
// with the sign of d * s (sign of d alone for s == 0, 0 for d == s == 0).  |d| / |s| is reduced to |t| <= tan(pi/16) with
// one of the base angles 0, pi/8, pi/4 (after swapping so that the ratio is <= 1): atan(a / b) = base + atan((a - c b) /
// (b + c a)), c = tan(base); ten Taylor terms then leave < 1e-16.  The base is picked from float32 copies of a and b (any
// choice near a threshold is fine).  Explicit fma() throughout: nothing here is pinned to numpy's rounding (scores are
// compared at 1e-9), unlike the float32 paths of this library.
__device__ __forceinline__ double rcp_lean(double x) {  // x > 0, normal
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  r = fma(r, fma(-x, r, 1.0), r);
  r = fma(r, fma(-x, r, 1.0), r);
  return r;
}
__device__ __forceinline__ double sqrt_lean(double x) {  // x >= 0
  double r;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  const double hx = 0.5 * x;
  r = fma(r, fma(-hx * r, r, 0.5), r);
  r = fma(r, fma(-hx * r, r, 0.5), r);
  return x > 0.0 ? x * r : 0.0;
}
__device__ __forceinline__ double abs_bits(double x) { return __hiloint2double(__double2hiint(x) & 0x7fffffff, __double2loint(x)); }
// d, s: the float64 difference and sum of the two channels; df, sf: float32 copies of them, only used to pick the
// reduction (which operand is larger, which base angle): any choice near a threshold is fine.
__device__ __forceinline__ double folded_angle(double d, double s, float df, float sf) {
  const double kC1 = 0.41421356237309503;  // tan(pi/8)
  const float fa = fabsf(df), fb = fabsf(sf);
  const bool inv = fa > fb;  // atan(a / b) = pi/2 - atan(b / a)
  const double ad = abs_bits(d), as = abs_bits(s);
  const double a = inv ? as : ad, b = inv ? ad : as;
  const float af = inv ? fb : fa, bf = inv ? fa : fb;
  const bool k1 = af > 0.19891237f * bf, k2 = af > 0.66817864f * bf;  // tan(pi/16), tan(3 pi/16)
  const double c = k2 ? 1.0 : (k1 ? kC1 : 0.0);
  const double base = k2 ? 0.7853981633974483 : (k1 ? 0.39269908169872414 : 0.0);
  const double num = fma(-c, b, a), den = fma(c, a, b);
  const double t = den > 0.0 ? num * rcp_lean(den) : 0.0;
  const double z = t * t;
  double p = -1.0 / 19.0;
  p = fma(p, z, 1.0 / 17.0);
  p = fma(p, z, -1.0 / 15.0);
  p = fma(p, z, 1.0 / 13.0);
  p = fma(p, z, -1.0 / 11.0);
  p = fma(p, z, 1.0 / 9.0);
  p = fma(p, z, -1.0 / 7.0);
  p = fma(p, z, 1.0 / 5.0);
  p = fma(p, z, -1.0 / 3.0);
  double th = base + fma(t * z, p, t);
  if (inv) th = 1.5707963267948966 - th;
  // sign of d * s (s == +-0 counts as positive: atan2(d, +-0) = sign(d) * pi/2)
  const int neg = (__double2hiint(d) ^ (s < 0.0 ? 0x80000000 : 0)) & 0x80000000;
  return __hiloint2double(__double2hiint(th) ^ neg, __double2loint(th));
}

struct HaasAcc {
  double v0 = 0, v1 = 0, v2 = 0, v3 = 0, v5 = 0, v6 = 0;
  long long vmax = 0;  // bits of max |theta| (non-negative doubles order like integers)
  __device__ __forceinline__ void add(double l, double r, float lf, float rf) {
    const double th = folded_angle(l - r, l + r, lf - rf, lf + rf);
    const double rad = sqrt_lean(fma(l, l, r * r));
    const double rt = rad * th, t2 = th * th;
    v0 += rad;
    v1 += rt;
    v2 = fma(rad, t2, v2);
    v3 = fma(rt, t2, v3);
    const long long ab = __double_as_longlong(th) & 0x7fffffffffffffffLL;
    vmax = ab > vmax ? ab : vmax;
    v5 = fma(l, r, v5);
    v6 = fma(l, l, v6);
  }
};

template <typename TIn>
__global__ void __launch_bounds__(256) haas_objective_kernel(const TIn* __restrict__ clips, long long frames, long long clip_stride,
                                                             long long chan_stride, const int* __restrict__ delays, int n_cand,
                                                             double* __restrict__ partials) {
  const int cand = blockIdx.x, clip = blockIdx.y;
  const int d = delays[cand];
  const TIn* __restrict__ x0 = clips + (long long)clip * clip_stride;
  const TIn* __restrict__ x1 = x0 + chan_stride;
  const long long total = frames + d;
  HaasAcc acc;
  // Three stretches (decorrelation.py:220-222: channel 0 delayed by d, both padded to frames + d): only channel 1 for
  // m < d, both channels for d <= m < frames, only the delayed channel 0 behind the end of channel 1.
  const long long head = d < frames ? d : frames;
  for (long long m = threadIdx.x; m < head; m += blockDim.x) {
    const TIn r = x1[m];
    acc.add(0.0, (double)r, 0.0f, (float)r);
  }
  if (d < frames) {
    const TIn* __restrict__ p0 = x0 + threadIdx.x;      // x0[m - d]
    const TIn* __restrict__ p1 = x1 + d + threadIdx.x;  // x1[m]
    const long long n = frames - d;
    for (long long k = threadIdx.x; k < n; k += blockDim.x, p0 += blockDim.x, p1 += blockDim.x) {
      const TIn l = *p0, r = *p1;
      acc.add((double)l, (double)r, (float)l, (float)r);
    }
  }  // (d >= frames: the frames between the two channels are silence and add nothing to any sum)
  const long long tail0 = frames > d ? frames : d;
  for (long long m = tail0 + threadIdx.x; m < total; m += blockDim.x) {
    const TIn l = x0[m - d];
    acc.add((double)l, 0.0, (float)l, 0.0f);
  }
  double v[7] = {acc.v0, acc.v1, acc.v2, acc.v3, __longlong_as_double(acc.vmax), acc.v5, acc.v6};
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
// The tensor-memory variant (vnd_objective_tmem.cu) is opt-in (VND_OBJ_TMEM=1): measured on 16 clips x 1024 strengths
// x 30 s it reaches 81 k evaluations/s against 136 k for the kernel below.  ncu (profiles/r02_summary.md): ~93 of the
// ~116 instructions per frame-evaluation are the polar moments, not the taps, so taking the taps off the shared-memory
// pipe (70 % busy here) buys little, while the 32-outputs-per-lane layout of the TMEM rows needs 163 registers: 12 warps
// per SM instead of 20 to hide the moments' dependent chains (issue 44 % instead of 63 %).
static bool obj_tmem_disabled() {
  const char* e = getenv("VND_OBJ_TMEM");
  return !(e && (e[0] == '1' || e[0] == '2'));
}
#define g_obj_disable_tmem obj_tmem_disabled()

struct ObjPlan {
  int cand_per_group, n_groups, tiles_per_chunk, n_chunks;
  size_t smem;
};

// Most chunks a clip's tiles are split into: enough for 16 waves of CTAs over all clips, never more than the tiles.
static long long max_chunks(long long tiles, int n_clips, int sm_count) {
  long long c = ceil_div<long long>((long long)sm_count * 16, n_clips > 0 ? n_clips : 1);
  if (c > tiles) c = tiles;
  return c < 1 ? 1 : c;
}

static int plan_objective(long long frames, int n_clips, int n_cand, int halo, int max_prog_words, ObjPlan* pl) {
  DeviceInfo di;
  int rc = device_info(&di);
  if (rc) return rc;
  const size_t span4 = (size_t)((OBJ_TILE + halo + 7) & ~3);
  const size_t fixed = (4 * span4 + OBJ_TILE) * 4 + (size_t)(OBJ_NT / 32) * 2 * max_prog_words * 4 + 64;
  // VND_OBJ_MINB CTAs share the SM's 228 KB (1 KB of each is reserved by the system); a single CTA may take it all
  size_t budget = (size_t)(233472 / VND_OBJ_MINB - 1024);
  if (budget > (size_t)kMaxDynSmem) budget = (size_t)kMaxDynSmem;
  if (fixed + OBJ_SLOTS * 8 * 16 > budget) budget = (size_t)kMaxDynSmem;  // long filters: one CTA per SM
  if (fixed + OBJ_SLOTS * 8 * 16 > budget) return VND_EUNSUPPORTED;
  int cpg = (int)((budget - fixed) / (OBJ_SLOTS * 8));
  if (cpg > n_cand) cpg = n_cand;
  if (cpg > 1024) cpg = 1024;
  const long long tiles = ceil_div<long long>(frames, OBJ_TILE);
  // One CTA per SM at a time: the launch takes ceil(CTAs / SMs) waves of (tiles per chunk) tile passes over the
  // candidates of a group.  First make the groups small enough for ~4 waves, then pick the split of the tiles into
  // chunks that minimises waves x tiles per chunk, i.e. that leaves the smallest idle tail in the last wave (64 clips
  // x 10 chunks = 640 CTAs on 148 SMs ran 5 waves for 4.3 waves of work; 37 chunks of 10 tiles run 16 full waves).
  const long long want = (long long)di.sm_count * 4;
  while ((long long)n_clips * tiles * ceil_div(n_cand, cpg) < want && cpg > 16) cpg = (cpg + 1) / 2;
  const long long groups = ceil_div(n_cand, cpg);
  long long n_chunks = 1, best = -1;
  const long long cap = max_chunks(tiles, n_clips, di.sm_count);  // what objective_workspace_bytes sizes the partial blocks for
  for (long long c = 1; c <= cap; ++c) {
    const long long tpc_c = ceil_div<long long>(tiles, c);
    const long long real_c = ceil_div<long long>(tiles, tpc_c);
    const long long waves = ceil_div<long long>((long long)n_clips * real_c * groups, di.sm_count);
    const long long cost = waves * tpc_c * 4096 + real_c;  // ties: fewer chunks (fewer partial blocks to combine)
    if (best < 0 || cost < best) {
      best = cost;
      n_chunks = real_c;
    }
  }
  int tpc = (int)ceil_div<long long>(tiles, n_chunks);
  if (const char* e = getenv("VND_OBJ_TPC")) {  // tuning: force the tiles per chunk (clamped to what the workspace was sized for)
    const int v = atoi(e);
    if (v > 0 && ceil_div<long long>(tiles, v) <= cap) tpc = v;
  }
  n_chunks = ceil_div<long long>(tiles, tpc);
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
  const long long chunks = max_chunks(tiles, n_clips, di.sm_count);
  *bytes = (size_t)n_clips * chunks * n_cand * OBJ_SLOTS * 8 + 256;
  const size_t tm_bytes = objective_tmem_workspace_bytes(frames, n_clips, n_cand, di.sm_count);  // the tensor-memory kernel tiles finer
  if (tm_bytes > *bytes) *bytes = tm_bytes;
  return VND_OK;
}

int vn_objective_launch(const float* clips, long long frames, int n_clips, long long clip_stride, long long chan_stride,
                        const vnd_tap_program* cand, double* partials, void* workspace, size_t workspace_bytes, cudaStream_t st) {
  if (n_clips == 0 || cand->channels == 0) return VND_OK;
  // the reference takes max(|theta|) of the frames: an empty signal is a ValueError there (optimization.py:41-43)
  VND_REQUIRE(frames > 0, VND_EINVAL, "objective of an empty signal (zero-size array to reduction operation maximum which has no identity)");
  if (!g_obj_disable_tmem) {  // taps through tensor memory whenever the family fits (vnd_objective_tmem.cu)
    const int rc_tm = vn_objective_tmem_launch(clips, frames, n_clips, clip_stride, chan_stride, cand, partials, workspace, workspace_bytes, st);
    if (rc_tm != VND_EUNSUPPORTED) return rc_tm;
  }
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
  return obj_combine_launch(p.chunk_partials, partials, n_clips, pl.n_chunks, cand->channels, st);
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
