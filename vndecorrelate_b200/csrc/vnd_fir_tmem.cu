// Tensor-memory sparse FIR (sm_100a): the throughput kernel for planar float32 slabs.
//
// Same arithmetic, in the same order, as VelvetNoise.convolve
// (src/vndecorrelate/decorrelation.py:393-415): bit-identical to fir_tile_kernel and the reference.
//
// Why tensor memory.  The FIR is a gather with data-dependent offsets: every (output, tap) pair
// moves one word from on-chip memory to a register.  Through shared memory that is 4 B x 30 taps
// per output against 128 B/clk/SM, which caps the kernel near 38 % of the HBM roofline however
// the loads are scheduled (profiles/r01_summary.md).  Blackwell has a second on-chip datapath
// into the register file: tcgen05.ld from tensor memory, measured at >= 390 B/clk/SM
// (tools/microbench/tmem_bw.cu) and independent of the LSU pipe.  tcgen05.ld.32x32b hands lane l
// of a warp N CONSECUTIVE columns of TMEM lane (row) l starting at a run-time column — exactly
// "R consecutive samples starting at a tap offset" if row l holds the samples that follow the
// first output of lane l.
//
// Layout.  A tile is 128 rows x R = 128 outputs.  TMEM row m (512 columns x 4 B) holds
// x[t0 + 128 m + c], c in [0, 512): a Hankel arrangement, each sample stored four times.  Thread
// (row m, group g) owns outputs 128 m + 32 g + r, r < 32; a tap with offset i <= 384 is one
// tcgen05.ld of 32 columns at column i + 32 g followed by 32 FADDs.  With log-distributed impulses
// 22 of the 30 taps of BASELINE config 3 qualify.  The remaining (far) taps read the staged tile
// from shared memory with 16-byte loads.
//
// Shared memory holds the tile as 128-sample blocks at a pitch of 132 words, so that lanes (rows
// 128 samples apart) hit distinct bank groups with LDS.128 / STS.128: one 512-byte TMA bulk copy
// per block.  Traffic on the LSU pipe per output: 4 words to fill TMEM (LDS.128 -> tcgen05.st),
// ~1.1 words per far tap, 1 word of output staging (STS.128 -> TMA bulk store), against ~27 for
// the register-window kernel.
//
// The kernel only runs interior tiles (tile + halo completely inside the signal); the launcher
// reports how many frames it covered and the caller finishes the tail of every channel with the
// general tile kernel.

#include "vnd_common.cuh"
#include "vnd_fir.cuh"

#ifndef VND_TM_RUN
#define VND_TM_RUN 32  // consecutive tiles of one channel per CTA run
#endif

namespace vnd {

namespace {

constexpr int kRows = 128;             // TMEM lanes = rows of a tile
constexpr int kR = 128;                // outputs per row
constexpr int kG = 4;                  // thread groups per row
constexpr int kRG = kR / kG;           // outputs per thread
constexpr int kNW = 4 * kG;            // warps per CTA (warp w: lane quarter w & 3, group w >> 2)
constexpr int kNT = kNW * 32;
constexpr int kTile = kRows * kR;      // 16384 outputs
constexpr int kPitch = kR + 4;         // block pitch in shared memory (words)
constexpr int kCols = 512;             // TMEM columns
constexpr int kNearMax = kCols - kR;   // largest tap offset served from TMEM
constexpr int kMaxBlocks = 156;        // 2 x 156 x 528 B + staging + program fits 227 KB
static_assert(kRG == 32, "a thread owns 32 outputs (two tcgen05.ld x16)");

// ---- tensor-memory primitives -------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc_all(uint32_t* slot) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(slot)) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_all(uint32_t addr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(addr) : "memory");
}
__device__ __forceinline__ void tmem_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

#define VND_O4(v, i) "=f"(v[i]), "=f"(v[i + 1]), "=f"(v[i + 2]), "=f"(v[i + 3])
#define VND_O16(v, i) VND_O4(v, i), VND_O4(v, i + 4), VND_O4(v, i + 8), VND_O4(v, i + 12)
#define VND_IO4(v, i) "+f"(v[i]), "+f"(v[i + 1]), "+f"(v[i + 2]), "+f"(v[i + 3])
#define VND_IO16(v, i) VND_IO4(v, i), VND_IO4(v, i + 4), VND_IO4(v, i + 8), VND_IO4(v, i + 12)
#define VND_I4(v, i) "f"(v[i].x), "f"(v[i].y), "f"(v[i].z), "f"(v[i].w)

// 16 consecutive columns of this thread's TMEM lane, starting at column (taddr & 0xffff).
template <int O>
__device__ __forceinline__ void tmem_ld16(float (&v)[kRG], uint32_t taddr) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : VND_O16(v, O)
               : "r"(taddr));
}
// The loaded registers are only defined after the wait; naming them as in/out operands keeps the
// compiler from moving their consumers above it.
__device__ __forceinline__ void tmem_wait_ld(float (&v)[kRG]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;" : VND_IO16(v, 0), VND_IO16(v, 16)::"memory");
}
// 32 consecutive columns written from eight float4.
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float4 (&v)[8]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%32], {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31};" ::VND_I4(v, 0),
      VND_I4(v, 1), VND_I4(v, 2), VND_I4(v, 3), VND_I4(v, 4), VND_I4(v, 5), VND_I4(v, 6), VND_I4(v, 7), "r"(taddr)
      : "memory");
}

struct TmParams {
  FirParams f;
  int nblk;  // 128-sample blocks staged per tile (tile + halo)
  int tiles_per_run;
  int runs_per_channel;
  long long n_runs;
  long long tiles_per_channel;  // interior tiles
};

// ---- packed fp32 arithmetic (sm_100+): FADD2 / FMUL2 do two IEEE round-to-nearest operations per
// instruction on an aligned register pair.  Same bits as two scalar operations, half the issue
// slots — and issue slots, not the FP32 pipe, are what this kernel runs out of.
typedef unsigned long long pair_t;
__device__ __forceinline__ pair_t pk(float a, float b) {
  pair_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void upk(pair_t r, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(r)); }
// In-place forms ("+l"): with a separate destination ptxas writes the result over the loaded operand
// and copies it back into the accumulator pair (32 extra moves per tap).
__device__ __forceinline__ pair_t add2(pair_t a, pair_t b) {
  asm("add.rn.f32x2 %0, %0, %1;" : "+l"(a) : "l"(b));
  return a;
}
__device__ __forceinline__ pair_t sub2(pair_t a, pair_t b) {
  asm("sub.rn.f32x2 %0, %0, %1;" : "+l"(a) : "l"(b));
  return a;
}
__device__ __forceinline__ pair_t mul2(pair_t a, pair_t b) {
  asm("mul.rn.f32x2 %0, %0, %1;" : "+l"(a) : "l"(b));
  return a;
}
constexpr int kNP = kRG / 2;  // register pairs per thread

__device__ __forceinline__ void bar_quarter(int q) { asm volatile("bar.sync %0, 128;" ::"r"(q + 1) : "memory"); }

// The negative list is accumulated with adds and negated once at the end of the list:
// fl(-a - b) == -fl(a + b) in round-to-nearest, and a zero of the other sign cannot survive into the
// output (the running output is never -0).  One add body per datapath.

// A tap served from tensor memory: acc += x[n + i] for the thread's 32 outputs.
__device__ __forceinline__ void near_tap(uint32_t tcol, pair_t (&acc)[kNP]) {
  float t[kRG];
  tmem_ld16<0>(t, tcol);
  tmem_ld16<16>(t, tcol + 16);
  tmem_wait_ld(t);
#pragma unroll
  for (int j = 0; j < kNP; ++j) acc[j] = add2(acc[j], pk(t[2 * j], t[2 * j + 1]));
}

// A tap served from shared memory.  `row` points at the staged block of this thread's row; `o` is
// the offset of the thread's first operand relative to it (32 g + i); A = o & 3 is warp-uniform.
// The 32 operands lie in 8 (A == 0) or 9 aligned 16-byte chunks; the run crosses at most one block
// boundary, where the pitch inserts a 4-word gap: KX = chunks before the gap.
template <int NC, int KX>
__device__ __forceinline__ void far_load(const float4* __restrict__ p, float4 (&c)[NC]) {
#pragma unroll
  for (int k = 0; k < NC; ++k) c[k] = p[k + (k >= KX ? 1 : 0)];
}

template <int A>
__device__ __forceinline__ void far_tap_a(const float* __restrict__ row, int o, pair_t (&acc)[kNP]) {
  constexpr int NC = (A == 0) ? 8 : 9;
  const int oal = o - A;
  const int w = oal & (kR - 1);
  const float4* p = reinterpret_cast<const float4*>(row + (oal >> 7) * kPitch + w);
  const int kx = (kR - w) >> 2;  // >= 1
  float4 c[NC];
  if (kx >= NC) {
    far_load<NC, NC>(p, c);
  } else {  // static addressing per crossing position
    switch (kx) {
      case 1: far_load<NC, 1>(p, c); break;
      case 2: far_load<NC, 2>(p, c); break;
      case 3: far_load<NC, 3>(p, c); break;
      case 4: far_load<NC, 4>(p, c); break;
      case 5: far_load<NC, 5>(p, c); break;
      case 6: far_load<NC, 6>(p, c); break;
      case 7: far_load<NC, 7>(p, c); break;
      default: far_load<NC, 8>(p, c); break;
    }
  }
  float t[NC * 4];
#pragma unroll
  for (int k = 0; k < NC; ++k) {
    t[4 * k] = c[k].x;
    t[4 * k + 1] = c[k].y;
    t[4 * k + 2] = c[k].z;
    t[4 * k + 3] = c[k].w;
  }
  if constexpr (A % 2 == 0) {  // operands arrive as aligned register pairs
#pragma unroll
    for (int j = 0; j < kNP; ++j)
      acc[j] = add2(acc[j], pk(t[2 * j + A], t[2 * j + 1 + A]));
  } else {  // odd shift: the pairs of the loaded data straddle the accumulator pairs -> scalar adds
#pragma unroll
    for (int j = 0; j < kNP; ++j) {
      float a0, a1;
      upk(acc[j], a0, a1);
      a0 = fadd(a0, t[2 * j + A]);
      a1 = fadd(a1, t[2 * j + 1 + A]);
      acc[j] = pk(a0, a1);
    }
  }
}

__device__ __forceinline__ void far_tap(const float* __restrict__ row, int o, pair_t (&acc)[kNP]) {
  switch (o & 3) {
    case 0: far_tap_a<0>(row, o, acc); break;
    case 1: far_tap_a<1>(row, o, acc); break;
    case 2: far_tap_a<2>(row, o, acc); break;
    default: far_tap_a<3>(row, o, acc); break;
  }
}

__device__ __forceinline__ void one_tap(int i, int og, uint32_t tcol0, const float* __restrict__ row, pair_t (&acc)[kNP]) {
  if (i <= kNearMax) near_tap(tcol0 + (uint32_t)i, acc);
  else far_tap(row, i + og, acc);
}

// Shared memory: [0,16) two mbarriers | [16,20) TMEM base | [64, ...) float in[2][nblk][132] |
//                float stage[128][132] | int program[]
__global__ void __launch_bounds__(kNT, 1) fir_tmem_kernel(const TmParams P) {
  const FirParams& p = P.f;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw);
  uint32_t* tm_slot = reinterpret_cast<uint32_t*>(smem_raw + 16);
  float* in_all = reinterpret_cast<float*>(smem_raw + 64);
  const int bufw = P.nblk * kPitch;
  float* stage = in_all + 2 * bufw;
  int* sprog = reinterpret_cast<int*>(stage + kRows * kPitch);

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int q = warp & 3, g = warp >> 2;
  const int m = 32 * q + lane;  // this thread's row (TMEM lane)

  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc_all(tm_slot);
  tmem_fence_before();
  __syncthreads();
  tmem_fence_after();
  const uint32_t tbase = *tm_slot + ((uint32_t)(32 * q) << 16);
  const uint32_t tcol0 = tbase + (uint32_t)(kRG * g);  // column of this thread's first output
  unsigned phases = 0;  // bit b: parity of the next completion of tile buffer b
  bool pending_store = false;

  for (int run = blockIdx.x; run < (int)P.n_runs; run += gridDim.x) {
    const int c = run / P.runs_per_channel;
    const int first_tile = (run % P.runs_per_channel) * P.tiles_per_run;
    int n_tiles = (int)P.tiles_per_channel - first_tile;
    if (n_tiles > P.tiles_per_run) n_tiles = P.tiles_per_run;
    const int w0 = p.offsets[c];
    const int nprog = p.offsets[c + 1] - w0;

    if (nprog == 0) {  // unfiltered channel: copy through (decorrelation.py:399-400)
      const float* xc = reinterpret_cast<const float*>(p.x) + (long long)c * p.x_sc;
      float* yc = p.y + (long long)c * p.y_sc;
      const long long t_begin = (long long)first_tile * kTile, t_end = t_begin + (long long)n_tiles * kTile;
      for (long long t = t_begin + 4 * tid; t < t_end; t += 4 * kNT)
        *reinterpret_cast<float4*>(yc + t) = *reinterpret_cast<const float4*>(xc + t);
      continue;
    }

    __syncthreads();  // everyone is done with the previous run's program
    for (int i = tid; i < nprog; i += kNT) sprog[i] = p.words[w0 + i];
    if (tid == 0) sprog[nprog] = 0;  // slack word read by the tap prefetch

    // one 512-byte bulk copy per 128-sample block.  UBLKCP is issued lane by lane, so the blocks are
    // dealt out to all warps (block b -> warp b % 16, lane b / 16): nine short issues per warp.
    auto issue = [&](int ti, int buf) {
      const float* src = reinterpret_cast<const float*>(p.x) + (long long)c * p.x_sc + (long long)(first_tile + ti) * kTile;
      float* dst = in_all + buf * bufw;
      if (tid == 0) mbar_expect_tx(&bars[buf], (uint32_t)P.nblk * (kR * 4u));
      const int b = warp + kNW * lane;
      if (b < P.nblk) {
        fence_proxy_async();
        bulk_g2s(dst + b * kPitch, src + b * kR, kR * 4u, &bars[buf]);
      }
    };
    issue(0, 0);
    __syncthreads();  // program visible
    const int S = sprog[0];

    for (int ti = 0; ti < n_tiles; ++ti) {
      const int buf = ti & 1;
      if (ti + 1 < n_tiles) issue(ti + 1, buf ^ 1);  // that buffer was released by the last barrier
      mbar_wait(&bars[buf], (phases >> buf) & 1u);
      phases ^= 1u << buf;
      const float* in = in_all + buf * bufw;
      const float* row = in + m * kPitch;

      // ---- fill: columns [128 g, 128 g + 128) of row m are block m + g ----
      {
        const float4* src = reinterpret_cast<const float4*>(in + (m + g) * kPitch);
#pragma unroll
        for (int k4 = 0; k4 < 4; ++k4) {
          float4 v[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] = src[8 * k4 + j];
          tmem_st32(tbase + (uint32_t)(kR * g + 32 * k4), v);
        }
        tmem_wait_st();
      }
      tmem_fence_before();
      bar_quarter(q);  // TMEM lanes are private to a lane quarter: its four warps are the only users
      tmem_fence_after();

      // ---- taps, in the reference's order: per segment the negative list, then the positive ----
      pair_t yv[kNP];
      {
        const int* seg = sprog + 1;
        const int* tp = sprog + 1 + 3 * S;
        for (int s = 0; s < S; ++s) {
          const int n_neg = seg[3 * s], n_pos = seg[3 * s + 1];
          pair_t acc[kNP];
#pragma unroll
          for (int j = 0; j < kNP; ++j) acc[j] = 0ull;
          int i_next = tp[0];  // one word of slack follows the program, so the prefetches stay in bounds
          const int n_tot = n_neg + n_pos;
          for (int k = 0; k < n_tot; ++k) {
            const int i = i_next;
            i_next = tp[k + 1];
            one_tap(i, kRG * g, tcol0, row, acc);
            if (k + 1 == n_neg) {  // end of the negative list: acc = -(sum of its taps)
              const pair_t m1 = pk(-1.0f, -1.0f);
#pragma unroll
              for (int j = 0; j < kNP; ++j) acc[j] = mul2(acc[j], m1);
            }
          }
          tp += n_tot;
          if (p.apply_gain) {
            const float gain = __int_as_float(seg[3 * s + 2]);
            const pair_t g2 = pk(gain, gain);
#pragma unroll
            for (int j = 0; j < kNP; ++j) acc[j] = mul2(acc[j], g2);
          }
          if (s == 0) {
#pragma unroll
            for (int j = 0; j < kNP; ++j) yv[j] = add2(acc[j], 0ull);  // the reference adds into zeros
          } else {
#pragma unroll
            for (int j = 0; j < kNP; ++j) yv[j] = add2(yv[j], acc[j]);
          }
        }
        if (S == 0) {
#pragma unroll
          for (int j = 0; j < kNP; ++j) yv[j] = 0ull;
        }
      }

      // ---- output: staging rows at pitch 132, then one 512-byte bulk store per row ----
      if (lane < 8 && pending_store) bulk_wait_read0();  // this warp's staging rows are free again
      tmem_fence_before();
      bar_quarter(q);  // every tcgen05.ld of this quarter is done (next fill), its staging rows are free
      tmem_fence_after();
      {
        float4* dst = reinterpret_cast<float4*>(stage + m * kPitch + kRG * g);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float4 v;
          upk(yv[2 * j], v.x, v.y);
          upk(yv[2 * j + 1], v.z, v.w);
          dst[j] = v;
        }
      }
      fence_proxy_async();
      bar_quarter(q);  // staging rows of this quarter are complete
      if (lane < 8) {  // row 32 q + 8 g + lane: eight 512-byte stores per warp
        const int sr = 32 * q + 8 * g + lane;
        float* yt = p.y + (long long)c * p.y_sc + (long long)(first_tile + ti) * kTile;
        bulk_s2g(yt + sr * kR, stage + sr * kPitch, kR * 4u);
        bulk_commit();
        pending_store = true;
      }
      __syncthreads();  // all reads of this tile buffer are done: it may be refilled
    }
  }
  if (lane < 8 && pending_store) bulk_wait_read0();  // shared memory must outlive the bulk reads
  tmem_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc_all(*tm_slot);
}

}  // namespace

// Runs the interior tiles of every channel and reports the frames covered per channel in
// *frames_done (a multiple of the tile).  VND_EUNSUPPORTED (no error text) when the request does
// not qualify; the caller then uses the other kernels for everything.
int fir_tmem_launch(const FirParams& f, int max_prog_words, cudaStream_t st, long long* frames_done) {
  *frames_done = 0;
  if (!f.bulk_ok || f.x_st != 1 || f.y_st != 1) return VND_EUNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(f.y) % 16) != 0 || (f.y_sc % 4) != 0) return VND_EUNSUPPORTED;
  const int nblk = kRows + (f.halo + 4 + kR - 1) / kR;
  if (nblk > kMaxBlocks) return VND_EUNSUPPORTED;
  if (f.frames < (long long)nblk * kR + 3LL * kTile) return VND_EUNSUPPORTED;  // fewer than four interior tiles
  const size_t smem = 64 + (size_t)2 * nblk * kPitch * 4 + (size_t)kRows * kPitch * 4 + (size_t)(max_prog_words + 4) * 4;
  if (smem > (size_t)kMaxDynSmem) return VND_EUNSUPPORTED;
  DeviceInfo di;
  int rc = device_info(&di);
  if (rc) return rc;
  TmParams P{};
  P.f = f;
  P.nblk = nblk;
  P.tiles_per_channel = (f.frames - (long long)nblk * kR) / kTile + 1;
  P.tiles_per_run = (int)(P.tiles_per_channel < VND_TM_RUN ? P.tiles_per_channel : VND_TM_RUN);
  P.runs_per_channel = (int)ceil_div<long long>(P.tiles_per_channel, P.tiles_per_run);
  P.n_runs = (long long)P.runs_per_channel * f.channels;
  VND_CUDA_OK(cudaFuncSetAttribute(fir_tmem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  long long grid = di.sm_count;
  if (grid > P.n_runs) grid = P.n_runs;
  fir_tmem_kernel<<<(unsigned)grid, kNT, smem, st>>>(P);
  rc = after_launch("fir_tmem_kernel");
  if (rc) return rc;
  *frames_done = P.tiles_per_channel * kTile;
  return VND_OK;
}

}  // namespace vnd
