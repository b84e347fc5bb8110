// Ring-buffer sparse FIR for long filters (sm_100a): BASELINE config 4 (300 impulses over 0.3 s @ 96 kHz, halo 28 800).
//
// Same arithmetic, in the same order, as VelvetNoise.convolve (src/vndecorrelate/decorrelation.py:393-415) and as
// fir_tile_kernel, whose run_program it shares in spirit: lane l of a warp owns outputs base + l + 32 r, a warp-uniform
// tap offset makes the 32 lanes read 32 consecutive words (conflict-free LDS.32, one word per (output, tap) pair - the
// floor for a gather through shared memory), packed adds.
//
// What is different is how the samples get there.  fir_tile_kernel stages tile + halo (16 384 + 28 452 samples, 177 KB)
// per CTA with one bulk copy and waits for it: the load is not overlapped with anything and every sample is fetched 2.7
// times (from L2).  Here one persistent CTA per SM walks along a channel in steps of C = 9216 outputs (12 warps x 32 lanes x 24) over a ring of
// N = ceil((C + halo) / C) + 1 chunks of C samples: a step needs chunks s .. s + K (K = N - 2), the chunk behind them is in
// flight while the step computes, and the chunk a step has left behind is the next one to be overwritten.  Every sample
// is fetched once, by a producer warp (one elected lane, cp.async.bulk + mbarrier), under the taps of the previous step.
//
// The kernel covers the interior of every channel (all chunks it touches lie inside the signal); the launcher reports
// how many frames that is and the caller finishes the tail of every channel with fir_tile_kernel.

#include <stdlib.h>

#include "vnd_common.cuh"
#include "vnd_fir.cuh"

namespace vnd {

namespace {

#ifndef VND_RING_R
#define VND_RING_R 24
#endif
// Shapes measured on BASELINE config 4 (128 channels x 57.6 M frames, Gsamples/s; fir_tile_kernel: 27.5): warps x outputs per
// lane 16 x 16: 28.0, 16 x 18: 28.2, 16 x 20: 28.3, 12 x 22: 28.3, 12 x 24: 28.6, 12 x 26: 28.7, 12 x 28: 28.6, 8 x 32: 26.6,
// 14 x 20: 26.1 and 10 x 28: 26.8 (warps not a multiple of the four schedulers), 20 x 16 with a producer warp: 24.9.
constexpr int kRingR = VND_RING_R;                // outputs per lane and step
#ifndef VND_RING_WARPS
#define VND_RING_WARPS 12
#endif
constexpr int kRingWarps = VND_RING_WARPS;        // compute warps (a multiple of four: one share per scheduler)
constexpr int kRingC = kRingWarps * 32 * kRingR;  // outputs per step = samples per chunk
constexpr int kRingNT = kRingWarps * 32;          // no producer warp: thread 0 feeds the ring between its own taps

struct RingParams {
  FirParams f;
  int n_chunks;         // N: ring slots
  int steps;            // steps per channel covered by this kernel
  int steps_per_run;
  int runs_per_channel;
  int n_runs;
};

typedef unsigned long long ring_pair_t;
#define VND_RING_PACKED(name, op)                                                  \
  __device__ __forceinline__ void name(float& a0, float& a1, float b0, float b1) { \
    ring_pair_t ra, rb;                                                            \
    asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(a0), "f"(a1));                   \
    asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(b0), "f"(b1));                   \
    asm(op ".rn.f32x2 %0, %0, %1;" : "+l"(ra) : "l"(rb));                         \
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a0), "=f"(a1) : "l"(ra));                  \
  }
VND_RING_PACKED(ring_add2, "add")
VND_RING_PACKED(ring_sub2, "sub")
VND_RING_PACKED(ring_mul2, "mul")

__device__ __forceinline__ bool ring_try(uint64_t* bar, uint32_t parity) {  // one poll of an mbarrier phase
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void ring_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// The kRingR operands of one tap for this lane: ring[pos + lane + 32 r].  `pos` (warp-uniform, < RB) is the ring offset of
// the operand of the lane-0 output; a warp's 32 * kRingR-sample span may run past the end of the ring into the mirror of
// the ring's first 32 * kRingR samples that the producer keeps behind it, so no lane ever wraps.
__device__ __forceinline__ void ring_gather(const float* __restrict__ ring, int pos, int lane, float (&t)[kRingR]) {
  const float* q = ring + pos + lane;
#pragma unroll
  for (int r = 0; r < kRingR; ++r) t[r] = q[32 * r];
}

// One tap of a stream: gather, then subtract (negative impulses come first) or add.
struct RingStream {
  const int* tp;  // tap offsets of the segment: negative impulses, then positive ones
  int n_neg, n;   // negative taps, all taps
  int nxt;        // offset of the next tap, loaded one tap ahead
};

// Dynamic shared memory: [0, 128) mbarriers full[N] (N <= 8) | [128, 256) mbarriers free[N] | float ring[N * C + 32 R] | int program[]
extern __shared__ __align__(128) unsigned char ring_smem[];

__global__ void __launch_bounds__(kRingNT, 1) fir_ring_kernel(const RingParams P) {
  const FirParams& p = P.f;
  uint64_t* full = reinterpret_cast<uint64_t*>(ring_smem);
  uint64_t* empty = reinterpret_cast<uint64_t*>(ring_smem + 128);
  float* ring = reinterpret_cast<float*>(ring_smem + 256);
  const int N = P.n_chunks, K = N - 2, RB = N * kRingC;
  int* sprog = reinterpret_cast<int*>(ring + RB + 32 * kRingR);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  for (int run = blockIdx.x; run < P.n_runs; run += gridDim.x) {
    const int c = run / P.runs_per_channel;
    const int s0 = (run % P.runs_per_channel) * P.steps_per_run;
    int ns = P.steps - s0;
    if (ns > P.steps_per_run) ns = P.steps_per_run;
    const int w0 = p.offsets[c];
    const int nprog = p.offsets[c + 1] - w0;
    const float* __restrict__ xc = reinterpret_cast<const float*>(p.x) + (long long)c * p.x_sc + (long long)s0 * kRingC;
    float* __restrict__ yc = p.y + (long long)c * p.y_sc + (long long)s0 * kRingC;
    if (nprog == 0) {  // unfiltered channel: copy through (decorrelation.py:399-400)
      const long long n = (long long)ns * kRingC;
      for (long long t = 4LL * tid; t < n; t += 4LL * kRingNT) *reinterpret_cast<float4*>(yc + t) = *reinterpret_cast<const float4*>(xc + t);
      continue;
    }
    __syncthreads();  // the previous run has drained: ring, barriers and program are free
    if (tid == 0) {
      for (int k = 0; k < N; ++k) {
        mbar_init(&full[k], 1);
        mbar_init(&empty[k], kRingWarps);
      }
      mbar_fence_init();
    }
    for (int i = tid; i < nprog; i += kRingNT) sprog[i] = p.words[w0 + i];
    __syncthreads();
    {
      // Thread 0 feeds the ring: it requests every chunk whose slot is free whenever it passes here (before a step and
      // between two pairs of segments), without ever blocking - a chunk is requested a whole step before it is needed.
      const int total = ns + K;  // chunks this run touches
      int jload = 0;
      auto pump = [&]() {
        if (tid != 0) return;
        while (jload < total) {
          const int slot = jload % N;
          if (jload >= N && !ring_try(&empty[slot], (unsigned)((jload / N - 1) & 1))) break;  // a warp still reads it
          const uint32_t mirror = slot == 0 ? 32u * kRingR * 4u : 0u;  // the head of slot 0 is kept twice: also behind the ring
          mbar_expect_tx(&full[slot], (uint32_t)kRingC * 4u + mirror);
          bulk_g2s(ring + (size_t)slot * kRingC, xc + (long long)jload * kRingC, (uint32_t)kRingC * 4u, &full[slot]);
          if (mirror) bulk_g2s(ring + RB, xc + (long long)jload * kRingC, mirror, &full[slot]);
          ++jload;
        }
      };
      pump();
      const int S = sprog[0];
      const int* seg = sprog + 1;
      for (int s = 0; s < ns; ++s) {
        pump();
        for (int j = (s == 0 ? 0 : s + K); j <= s + K; ++j) {
          // Thread 0 is the only producer: while it waits for a chunk it keeps feeding the ring (the chunk it waits for may
          // not even be requested yet, if another warp was still reading the slot when thread 0 last looked); the other
          // lanes spin.  Waiting with a plain spin in thread 0 deadlocks as soon as a warp lags by more than a step.
          if (tid == 0) {
            while (!ring_try(&full[j % N], (unsigned)((j / N) & 1))) pump();
          } else {
            mbar_wait(&full[j % N], (unsigned)((j / N) & 1));
          }
        }
        const int pos0 = (s % N) * kRingC + 32 * kRingR * warp;  // ring offset of this warp's first output sample
        float yv[kRingR];
#pragma unroll
        for (int r = 0; r < kRingR; ++r) yv[r] = 0.0f;
        // Two decay segments at a time: their sums are independent (decorrelation.py:402-414: `acc` restarts per segment),
        // so their taps are interleaved - the loads of both are issued before either's adds - which doubles the loads a
        // warp has in flight (16 warps per SM have to keep the shared-memory pipe busy).  The running output still adds the
        // scaled sums in segment order.
        const int* tp = sprog + 1 + 3 * S;
        for (int sg = 0; sg < S; sg += 2) {
          if (sg) pump();
          const bool two = sg + 1 < S;
          RingStream A{tp, seg[3 * sg + 0], seg[3 * sg + 0] + seg[3 * sg + 1], 0};
          tp += A.n;
          RingStream B{tp, two ? seg[3 * sg + 3] : 0, two ? seg[3 * sg + 3] + seg[3 * sg + 4] : 0, 0};
          tp += B.n;
          float accA[kRingR], accB[kRingR];
#pragma unroll
          for (int r = 0; r < kRingR; ++r) {
            accA[r] = 0.0f;
            accB[r] = 0.0f;
          }
          A.nxt = A.tp[0];
          B.nxt = B.tp[0];  // (one word of slack follows the program)
          const int nboth = A.n < B.n ? A.n : B.n;
          int k = 0;
          for (; k < nboth; ++k) {
            int pa = pos0 + A.nxt, pb = pos0 + B.nxt;
            A.nxt = A.tp[k + 1];
            B.nxt = B.tp[k + 1];
            if (pa >= RB) pa -= RB;
            if (pb >= RB) pb -= RB;
            float ta[kRingR], tb[kRingR];
            ring_gather(ring, pa, lane, ta);
            ring_gather(ring, pb, lane, tb);
            if (k < A.n_neg) {
#pragma unroll
              for (int r = 0; r < kRingR; r += 2) ring_sub2(accA[r], accA[r + 1], ta[r], ta[r + 1]);
            } else {
#pragma unroll
              for (int r = 0; r < kRingR; r += 2) ring_add2(accA[r], accA[r + 1], ta[r], ta[r + 1]);
            }
            if (k < B.n_neg) {
#pragma unroll
              for (int r = 0; r < kRingR; r += 2) ring_sub2(accB[r], accB[r + 1], tb[r], tb[r + 1]);
            } else {
#pragma unroll
              for (int r = 0; r < kRingR; r += 2) ring_add2(accB[r], accB[r + 1], tb[r], tb[r + 1]);
            }
          }
          for (int ka = k; ka < A.n; ++ka) {  // the longer segment alone
            int pa = pos0 + A.tp[ka];
            if (pa >= RB) pa -= RB;
            float ta[kRingR];
            ring_gather(ring, pa, lane, ta);
            if (ka < A.n_neg) {
#pragma unroll
              for (int r = 0; r < kRingR; r += 2) ring_sub2(accA[r], accA[r + 1], ta[r], ta[r + 1]);
            } else {
#pragma unroll
              for (int r = 0; r < kRingR; r += 2) ring_add2(accA[r], accA[r + 1], ta[r], ta[r + 1]);
            }
          }
          for (int kb = k; kb < B.n; ++kb) {
            int pb = pos0 + B.tp[kb];
            if (pb >= RB) pb -= RB;
            float tb[kRingR];
            ring_gather(ring, pb, lane, tb);
            if (kb < B.n_neg) {
#pragma unroll
              for (int r = 0; r < kRingR; r += 2) ring_sub2(accB[r], accB[r + 1], tb[r], tb[r + 1]);
            } else {
#pragma unroll
              for (int r = 0; r < kRingR; r += 2) ring_add2(accB[r], accB[r + 1], tb[r], tb[r + 1]);
            }
          }
          if (p.apply_gain) {
            const float ga = __int_as_float(seg[3 * sg + 2]);
#pragma unroll
            for (int r = 0; r < kRingR; r += 2) ring_mul2(accA[r], accA[r + 1], ga, ga);
          }
#pragma unroll
          for (int r = 0; r < kRingR; r += 2) ring_add2(yv[r], yv[r + 1], accA[r], accA[r + 1]);
          if (two) {
            if (p.apply_gain) {
              const float gb = __int_as_float(seg[3 * sg + 5]);
#pragma unroll
              for (int r = 0; r < kRingR; r += 2) ring_mul2(accB[r], accB[r + 1], gb, gb);
            }
#pragma unroll
            for (int r = 0; r < kRingR; r += 2) ring_add2(yv[r], yv[r + 1], accB[r], accB[r + 1]);
          }
        }
        __syncwarp();
        if (lane == 0) ring_arrive(&empty[s % N]);  // chunk s is behind this warp
        float* yo = yc + (long long)s * kRingC + 32 * kRingR * warp + lane;
#pragma unroll
        for (int r = 0; r < kRingR; ++r) yo[32 * r] = yv[r];
      }
    }
  }
}

}  // namespace

static const bool g_disable_ring = [] {
  const char* e = getenv("VND_DISABLE_RING");
  return e && e[0] == '1';
}();

// Runs the interior of every channel and reports the frames covered per channel in *frames_done (a multiple of the
// step).  VND_EUNSUPPORTED (no error text) when the request does not qualify; the caller then uses fir_tile_kernel for
// everything.
int fir_ring_launch(const FirParams& f, int max_prog_words, cudaStream_t st, long long* frames_done) {
  *frames_done = 0;
  if (g_disable_ring || !f.bulk_ok || f.x_st != 1 || f.y_st != 1) return VND_EUNSUPPORTED;
  if (f.channels > 1 && (f.x_sc % 4 != 0)) return VND_EUNSUPPORTED;
  const int K = (kRingC + f.halo + kRingC - 1) / kRingC - 1;  // a step reads chunks s .. s + K
  const int N = K + 2;
  if (N > 8) return VND_EUNSUPPORTED;
  const size_t smem = 256 + ((size_t)N * kRingC + 32 * kRingR) * 4 + (size_t)max_prog_words * 4 + 16;
  if (smem > (size_t)kMaxDynSmem) return VND_EUNSUPPORTED;
  // the max tap offset must stay below (K + 1) * C - C = K * C + ... : guaranteed by K; the wrap logic needs halo + C < RB
  const long long steps = f.frames / kRingC - K;  // every chunk a step touches lies inside the signal
  if (steps < 8) return VND_EUNSUPPORTED;
  if (steps > 0x3fffffffLL / (f.channels > 0 ? f.channels : 1)) return VND_EUNSUPPORTED;
  DeviceInfo di;
  int rc = device_info(&di);
  if (rc) return rc;
  RingParams P{};
  P.f = f;
  P.n_chunks = N;
  P.steps = (int)steps;
  // equal runs so that the persistent CTAs finish together: fewest (waves of runs) x (steps per run + pipeline fill)
  long long best_cost = -1;
  int best_rpc = 1;
  for (int rpc = 1; rpc <= 512 && rpc <= steps; ++rpc) {
    const long long spr = ceil_div<long long>(steps, rpc);
    const long long real_rpc = ceil_div<long long>(steps, spr);
    const long long waves = ceil_div<long long>(real_rpc * f.channels, di.sm_count);
    const long long cost = waves * (spr + 1);
    if (best_cost < 0 || cost < best_cost) {
      best_cost = cost;
      best_rpc = (int)real_rpc;
    }
  }
  P.steps_per_run = (int)ceil_div<long long>(steps, best_rpc);
  P.runs_per_channel = (int)ceil_div<long long>(steps, P.steps_per_run);
  P.n_runs = P.runs_per_channel * f.channels;
  VND_CUDA_OK(cudaFuncSetAttribute(fir_ring_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int grid = di.sm_count;
  if (grid > P.n_runs) grid = P.n_runs;
  fir_ring_kernel<<<(unsigned)grid, kRingNT, smem, st>>>(P);
  rc = after_launch("fir_ring_kernel");
  if (rc) return rc;
  *frames_done = steps * kRingC;
  return VND_OK;
}

}  // namespace vnd
