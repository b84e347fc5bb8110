// Microbenchmark (sm_100a): what limits the tensor-memory taps of fir_tmem_kernel?
//
// Every warp runs the tap loop of the kernel on its own TMEM lane quarter: a warp-uniform,
// data-dependent column offset read from shared memory, one tcgen05.ld.32x32b.xN at that column,
// N/2 packed adds (FADD2) into N accumulators.  Variants:
//   single   ld -> wait -> adds                         (the kernel's loop: one buffer)
//   pipe     ld(k+1) issued before the adds of tap k    (two buffers, one load in flight behind the adds)
//   pipe3    two loads in flight behind the adds        (three buffers)
//   half     the x32 tap as two x16 loads, the second one in flight behind the first one's adds
// swept over warps per SM (NW: 4 = one per scheduler ... 16 = four per scheduler) and over the
// alignment of the column offsets (any / multiple of 4 / of 32).  Output: clocks per tap and warp,
// bytes per clock and SM.
//
// Build: nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -o tmem_lat tmem_lat.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x)                                                                        \
  do {                                                                               \
    cudaError_t e = (x);                                                             \
    if (e != cudaSuccess) {                                                          \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); \
      exit(1);                                                                       \
    }                                                                                \
  } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

#define O4(v, i) "=f"(v[i]), "=f"(v[i + 1]), "=f"(v[i + 2]), "=f"(v[i + 3])
#define O16(v, i) O4(v, i), O4(v, i + 4), O4(v, i + 8), O4(v, i + 12)
#define IO4(v, i) "+f"(v[i]), "+f"(v[i + 1]), "+f"(v[i + 2]), "+f"(v[i + 3])
#define IO16(v, i) IO4(v, i), IO4(v, i + 4), IO4(v, i + 8), IO4(v, i + 12)

template <int N>
__device__ __forceinline__ void ldtm(float (&v)[N], uint32_t taddr);
template <>
__device__ __forceinline__ void ldtm<16>(float (&v)[16], uint32_t taddr) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];" : O16(v, 0) : "r"(taddr));
}
template <>
__device__ __forceinline__ void ldtm<32>(float (&v)[32], uint32_t taddr) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : O16(v, 0), O16(v, 16)
      : "r"(taddr));
}
template <int N>
__device__ __forceinline__ void wait_ld(float (&v)[N]);
template <>
__device__ __forceinline__ void wait_ld<16>(float (&v)[16]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;" : IO16(v, 0)::"memory");
}
template <>
__device__ __forceinline__ void wait_ld<32>(float (&v)[32]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;" : IO16(v, 0), IO16(v, 16)::"memory");
}
__device__ __forceinline__ void add2(float& a0, float& a1, float b0, float b1) {
  unsigned long long ra, rb;
  asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(a0), "f"(a1));
  asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(b0), "f"(b1));
  asm("add.rn.f32x2 %0, %0, %1;" : "+l"(ra) : "l"(rb));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a0), "=f"(a1) : "l"(ra));
}
template <int N, int O = 0>
__device__ __forceinline__ void adds(float (&acc)[32], const float (&t)[N]) {
#pragma unroll
  for (int j = 0; j < N; j += 2) add2(acc[O + j], acc[O + j + 1], t[j], t[j + 1]);
}

enum Mode { SINGLE = 0, PIPE = 1, PIPE3 = 2, HALF = 3 };

struct Result {
  unsigned long long cycles;
  float sink;
};

template <int MODE, int NW>
__global__ void __launch_bounds__(NW * 32, 1) bench_kernel(int iters, int align, Result* out) {
  __shared__ int offs[64 + 8];
  __shared__ uint32_t tm_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid < 72) offs[tid] = ((tid * 37 + 11) % 416) & ~(align - 1);
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tm_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tbase = tm_slot + ((uint32_t)((warp & 3) * 32) << 16) + 32 * (warp >> 2) % 96;
  float acc[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) acc[j] = 0.0f;
  __syncthreads();
  const unsigned long long t0 = clock64();
  if constexpr (MODE == SINGLE) {
    int o = offs[0];
    for (int it = 0; it < iters; ++it) {
      const int on = offs[(it + 1) & 63];
      float t[32];
      ldtm<32>(t, tbase + o);
      wait_ld<32>(t);
      adds<32>(acc, t);
      o = on;
    }
  } else if constexpr (MODE == PIPE) {
    float A[32], B[32];
    ldtm<32>(A, tbase + offs[0]);
    int o1 = offs[1];
    for (int it = 0; it < iters; it += 2) {
      const int o2 = offs[(it + 2) & 63];
      wait_ld<32>(A);
      ldtm<32>(B, tbase + o1);
      adds<32>(acc, A);
      o1 = offs[(it + 3) & 63];
      wait_ld<32>(B);
      ldtm<32>(A, tbase + o2);
      adds<32>(acc, B);
    }
    wait_ld<32>(A);
    adds<32>(acc, A);
  } else if constexpr (MODE == HALF) {
    float A[16], B[16];
    int o = offs[0];
    ldtm<16>(A, tbase + o);
    for (int it = 0; it < iters; ++it) {
      const int on = offs[(it + 1) & 63];
      wait_ld<16>(A);
      ldtm<16>(B, tbase + o + 16);
      adds<16, 0>(acc, A);
      wait_ld<16>(B);
      ldtm<16>(A, tbase + on);
      adds<16, 16>(acc, B);
      o = on;
    }
    wait_ld<16>(A);
    adds<16, 0>(acc, A);
  } else {  // PIPE3: only 16 accumulators are live per tap half, so that three x32 buffers fit 4 warps per scheduler
    float A[32], B[32], C[32];
    ldtm<32>(A, tbase + offs[0]);
    ldtm<32>(B, tbase + offs[1]);
    for (int it = 0; it < iters; it += 3) {
      wait_ld<32>(A);  // waits for B too (tcgen05.wait::ld has no finer grain)
      ldtm<32>(C, tbase + offs[(it + 2) & 63]);
      adds<32>(acc, A);
      ldtm<32>(A, tbase + offs[(it + 3) & 63]);
      adds<32>(acc, B);
      wait_ld<32>(C);
      ldtm<32>(B, tbase + offs[(it + 4) & 63]);
      adds<32>(acc, C);
    }
    wait_ld<32>(A);
    adds<32>(acc, A);
    adds<32>(acc, B);
  }
  __syncthreads();
  const unsigned long long t1 = clock64();
  float s = 0.0f;
#pragma unroll
  for (int j = 0; j < 32; ++j) s += acc[j];
  if (tid == 0) out[blockIdx.x].cycles = t1 - t0;
  if (s == 123.456f) out[blockIdx.x].sink = s;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm_slot) : "memory");
}

// Interference between lane quarters: the warps of quarters < QT run the single-buffer tensor-memory tap
// loop, the warps of the other quarters run the kernel's shared-memory tap (nine 16-byte loads from a
// conflict-free row layout + 16 packed adds).  Each group reports its own clocks.
template <int NW, int QT>
__global__ void __launch_bounds__(NW * 32, 1) mixq_kernel(int iters, Result* out_t, Result* out_s) {
  __shared__ int offs[64 + 8];
  __shared__ uint32_t tm_slot;
  extern __shared__ __align__(16) float tile[];  // 128 rows x 100 words + slack
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid < 72) offs[tid] = ((tid * 37 + 11) % 416);
  for (int i = tid; i < 128 * 100 + 1024; i += NW * 32) tile[i] = (float)i;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tm_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const int q = warp & 3;
  const uint32_t tbase = tm_slot + ((uint32_t)(q * 32) << 16) + 32 * (warp >> 2) % 96;
  float acc[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) acc[j] = 0.0f;
  __syncthreads();
  const unsigned long long t0 = clock64();
  if (q < QT) {
    int o = offs[0];
    for (int it = 0; it < iters; ++it) {
      const int on = offs[(it + 1) & 63];
      float t[32];
      ldtm<32>(t, tbase + o);
      wait_ld<32>(t);
      adds<32>(acc, t);
      o = on;
    }
  } else {
    const float* row = tile + (32 * q + lane) * 100;
    int o = offs[0];
    for (int it = 0; it < iters; ++it) {
      const int on = offs[(it + 1) & 63];
      const float4* p = reinterpret_cast<const float4*>(row + (o & ~3));
      float4 c[9];
#pragma unroll
      for (int k = 0; k < 9; ++k) c[k] = p[k];
      float t[32];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        t[4 * k] = c[k].x; t[4 * k + 1] = c[k].y; t[4 * k + 2] = c[k].z; t[4 * k + 3] = c[k].w;
      }
      adds<32>(acc, t);
      acc[0] += c[8].x;
      o = on;
    }
  }
  const unsigned long long t1 = clock64();
  float s = 0.0f;
#pragma unroll
  for (int j = 0; j < 32; ++j) s += acc[j];
  if (lane == 0 && warp == 0) out_t[blockIdx.x].cycles = t1 - t0;
  if (lane == 0 && warp == 3) out_s[blockIdx.x].cycles = t1 - t0;
  if (s == 123.456f) out_t[blockIdx.x].sink = s;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm_slot) : "memory");
}

// Interference of the TMEM refill: 12 warps run the single-buffer tensor-memory tap loop (or the
// pipelined one, PIPED); warps 12..15 (one per lane quarter) refill their quarter's rows the way the
// kernel's data-movement warps do (8 x LDS.128 -> tcgen05.st.x32, 16 per row) when their quarter is
// < QF, in a loop with DUTY % duty (sleeping in between).  Columns 0..415 are written, the taps
// read the same columns: values are garbage, only the timing matters.
template <int QF, bool PIPED>
__global__ void __launch_bounds__(512, 1) fill_kernel(int iters, int sleep_ns, Result* out_t, Result* out_s) {
  __shared__ int offs[64 + 8];
  __shared__ uint32_t tm_slot;
  __shared__ int done;
  extern __shared__ __align__(16) float tile[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid < 72) offs[tid] = ((tid * 37 + 11) % 416);
  if (tid == 0) done = 0;
  for (int i = tid; i < 128 * 100 + 1024; i += 512) tile[i] = (float)i;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tm_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const int q = warp & 3;
  const uint32_t tlane = tm_slot + ((uint32_t)(q * 32) << 16);
  const unsigned long long t0 = clock64();
  if (warp < 12) {
    const uint32_t tbase = tlane + 32 * (warp >> 2);
    float acc[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) acc[j] = 0.0f;
    if constexpr (!PIPED) {
      int o = offs[0];
      for (int it = 0; it < iters; ++it) {
        const int on = offs[(it + 1) & 63];
        float t[32];
        ldtm<32>(t, tbase + o);
        wait_ld<32>(t);
        adds<32>(acc, t);
        o = on;
      }
    } else {
      float A[32], B[32];
      ldtm<32>(A, tbase + offs[0]);
      int o1 = offs[1];
      for (int it = 0; it < iters; it += 2) {
        const int o2 = offs[(it + 2) & 63];
        wait_ld<32>(A);
        ldtm<32>(B, tbase + o1);
        adds<32>(acc, A);
        o1 = offs[(it + 3) & 63];
        wait_ld<32>(B);
        ldtm<32>(A, tbase + o2);
        adds<32>(acc, B);
      }
      wait_ld<32>(A);
      adds<32>(acc, A);
    }
    const unsigned long long t1 = clock64();
    float s = 0.0f;
#pragma unroll
    for (int j = 0; j < 32; ++j) s += acc[j];
    if (lane == 0 && warp == 0) out_t[blockIdx.x].cycles = t1 - t0;
    if (lane == 0 && warp == 3) out_s[blockIdx.x].cycles = t1 - t0;
    if (s == 123.456f) out_t[blockIdx.x].sink = s;
    if (lane == 0) atomicAdd(&done, 1);
  } else if (q < QF) {
    const float4* src0 = reinterpret_cast<const float4*>(tile + (32 * q + lane) * 100);
    while (*(volatile int*)&done < 12) {
      const float4* src = src0;
      uint32_t tcol = tlane;
      for (int k32 = 0; k32 < 13; ++k32) {
        float4 v[8];
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) v[jj] = src[jj];
        asm volatile(
            "tcgen05.st.sync.aligned.32x32b.x32.b32 [%32], {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
            "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31};" ::"f"(v[0].x),
            "f"(v[0].y), "f"(v[0].z), "f"(v[0].w), "f"(v[1].x), "f"(v[1].y), "f"(v[1].z), "f"(v[1].w), "f"(v[2].x), "f"(v[2].y), "f"(v[2].z), "f"(v[2].w),
            "f"(v[3].x), "f"(v[3].y), "f"(v[3].z), "f"(v[3].w), "f"(v[4].x), "f"(v[4].y), "f"(v[4].z), "f"(v[4].w), "f"(v[5].x), "f"(v[5].y), "f"(v[5].z),
            "f"(v[5].w), "f"(v[6].x), "f"(v[6].y), "f"(v[6].z), "f"(v[6].w), "f"(v[7].x), "f"(v[7].y), "f"(v[7].z), "f"(v[7].w), "r"(tcol)
            : "memory");
        tcol += 32;
        src += 8;
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      if (sleep_ns) __nanosleep(sleep_ns);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm_slot) : "memory");
}

template <int QF, bool PIPED>
void run_fill(int iters, int sleep_ns, Result* d_out, int sms) {
  auto k = fill_kernel<QF, PIPED>;
  const size_t smem = (128 * 100 + 1024) * 4;
  CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  for (int rep = 0; rep < 2; ++rep) {
    k<<<sms, 512, smem>>>(iters, sleep_ns, d_out, d_out + sms);
    CK(cudaDeviceSynchronize());
  }
  Result* h = (Result*)malloc(sizeof(Result) * 2 * sms);
  CK(cudaMemcpy(h, d_out, sizeof(Result) * 2 * sms, cudaMemcpyDeviceToHost));
  double ct = 0, cs = 0;
  for (int i = 0; i < sms; ++i) {
    ct += (double)h[i].cycles;
    cs += (double)h[sms + i].cycles;
  }
  printf("fill    %s taps, refilling quarters=%d sleep=%d ns: tap of quarter 0 %.1f clk, tap of quarter 3 %.1f clk (per tap and warp)\n", PIPED ? "pipelined" : "single   ", QF,
         sleep_ns, ct / sms / iters, cs / sms / iters);
  free(h);
}

template <int NW, int QT>
void run_mixq(int iters, Result* d_out, int sms) {
  auto k = mixq_kernel<NW, QT>;
  const size_t smem = (128 * 100 + 1024) * 4;
  CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  for (int rep = 0; rep < 2; ++rep) {
    k<<<sms, NW * 32, smem>>>(iters, d_out, d_out + sms);
    CK(cudaDeviceSynchronize());
  }
  Result* h = (Result*)malloc(sizeof(Result) * 2 * sms);
  CK(cudaMemcpy(h, d_out, sizeof(Result) * 2 * sms, cudaMemcpyDeviceToHost));
  double ct = 0, cs = 0;
  for (int i = 0; i < sms; ++i) {
    ct += (double)h[i].cycles;
    cs += (double)h[sms + i].cycles;
  }
  printf("mixq    NW=%2d tensor-memory quarters=%d  TMEM tap %.1f clk per tap and warp   LDS.128 tap %.1f clk per tap and warp\n", NW, QT, ct / sms / iters,
         QT < 4 ? cs / sms / iters : 0.0);
  free(h);
}

template <int MODE, int NW>
void run(const char* name, int iters, int align, Result* d_out, int sms) {
  auto k = bench_kernel<MODE, NW>;
  k<<<sms, NW * 32>>>(iters, align, d_out);
  CK(cudaDeviceSynchronize());
  k<<<sms, NW * 32>>>(iters, align, d_out);
  CK(cudaDeviceSynchronize());
  Result* h = (Result*)malloc(sizeof(Result) * sms);
  CK(cudaMemcpy(h, d_out, sizeof(Result) * sms, cudaMemcpyDeviceToHost));
  double cyc = 0;
  for (int i = 0; i < sms; ++i) cyc += (double)h[i].cycles;
  cyc /= sms;
  printf("%-7s NW=%2d align=%2d  %.1f clk per tap and warp  %.1f B/clk/SM\n", name, NW, align, cyc / iters, (double)NW * 32 * 32 * 4 * iters / cyc);
  free(h);
}

template <int NW>
void all(int iters, Result* d_out, int sms) {
  for (int align : {1, 4, 32}) {
    run<SINGLE, NW>("single", iters, align, d_out, sms);
    run<PIPE, NW>("pipe", iters, align, d_out, sms);
    run<HALF, NW>("half", iters, align, d_out, sms);
    if (NW <= 12) run<PIPE3, NW>("pipe3", iters, align, d_out, sms);
  }
}

int main() {
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  printf("device %s, %d SMs\n", prop.name, sms);
  Result* d_out;
  CK(cudaMalloc(&d_out, sizeof(Result) * 2 * sms));
  const int iters = 6144;
  run_mixq<12, 4>(iters, d_out, sms);
  run_mixq<12, 3>(iters, d_out, sms);
  run_mixq<12, 2>(iters, d_out, sms);
  run_mixq<12, 1>(iters, d_out, sms);
  run_mixq<12, 0>(iters, d_out, sms);
  for (int sl : {0, 2000}) {
    run_fill<0, false>(iters, sl, d_out, sms);
    run_fill<1, false>(iters, sl, d_out, sms);
    run_fill<4, false>(iters, sl, d_out, sms);
    run_fill<0, true>(iters, sl, d_out, sms);
    run_fill<1, true>(iters, sl, d_out, sms);
    run_fill<4, true>(iters, sl, d_out, sms);
  }
  if (getenv("MIXQ_ONLY")) return 0;
  all<4>(iters, d_out, sms);
  all<8>(iters, d_out, sms);
  all<12>(iters, d_out, sms);
  all<16>(iters, d_out, sms);
  return 0;
}
