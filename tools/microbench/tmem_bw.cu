// Microbenchmark (sm_100a): is tensor memory a usable second gather datapath for the sparse FIR?
//
// Measures, per SM, with one 512-column allocation per CTA and NW warps:
//   ldtm   tcgen05.ld.32x32b.xN at warp-uniform, data-dependent column offsets, each loaded word
//          consumed by one FADD (what a FIR tap does)
//   lds    the same loop served by conflict-free scalar LDS (the current kernel's datapath)
//   mixed  one LDTM.xN tap and one LDS tap alternating (are the two pipes independent?)
//   sttm   tcgen05.st.32x32b.xN (the fill path from registers)
// and checks that a row written with tcgen05.st reads back correctly from an arbitrary (unaligned)
// column offset.  Output: bytes per clock per SM for each (mode, NW, N).
//
// Build: nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -o tmem_bw tmem_bw.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x)                                                                          \
  do {                                                                                 \
    cudaError_t e = (x);                                                               \
    if (e != cudaSuccess) {                                                            \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__);   \
      exit(1);                                                                         \
    }                                                                                  \
  } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void tmem_alloc512(uint32_t* slot) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(slot)) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc512(uint32_t addr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(addr) : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

#define R4(v, i) "=r"(v[i]), "=r"(v[i + 1]), "=r"(v[i + 2]), "=r"(v[i + 3])
#define R16(v, i) R4(v, i), R4(v, i + 4), R4(v, i + 8), R4(v, i + 12)
#define W4(v, i) "r"(v[i]), "r"(v[i + 1]), "r"(v[i + 2]), "r"(v[i + 3])
#define W16(v, i) W4(v, i), W4(v, i + 4), W4(v, i + 8), W4(v, i + 12)

__device__ __forceinline__ void ldtm16(uint32_t (&v)[16], uint32_t taddr) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : R16(v, 0)
      : "r"(taddr));
}
__device__ __forceinline__ void ldtm32(uint32_t (&v)[32], uint32_t taddr) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : R16(v, 0), R16(v, 16)
      : "r"(taddr));
}
__device__ __forceinline__ void ldtm64(uint32_t (&v)[64], uint32_t taddr) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x64.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,"
      "%32,%33,%34,%35,%36,%37,%38,%39,%40,%41,%42,%43,%44,%45,%46,%47,"
      "%48,%49,%50,%51,%52,%53,%54,%55,%56,%57,%58,%59,%60,%61,%62,%63}, [%64];"
      : R16(v, 0), R16(v, 16), R16(v, 32), R16(v, 48)
      : "r"(taddr));
}
__device__ __forceinline__ void sttm32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%32], {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31};" ::W16(v, 0),
      W16(v, 16), "r"(taddr)
      : "memory");
}

template <int N>
__device__ __forceinline__ void ldtm(uint32_t (&v)[N], uint32_t taddr) {
  if constexpr (N == 16) ldtm16(v, taddr);
  else if constexpr (N == 32) ldtm32(v, taddr);
  else ldtm64(v, taddr);
}

enum Mode { LDTM = 0, LDS = 1, MIXED = 2, STTM = 3, LDTM_NOWAIT2 = 4, FADD1 = 5, FADD2 = 6, LDTM_FADD2 = 7 };

// packed fp32 add/sub (sm_100+): two IEEE round-to-nearest adds per instruction (FADD2 in SASS)
__device__ __forceinline__ void add2(float& a0, float& a1, float b0, float b1) {
  unsigned long long ra, rb;
  asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(a0), "f"(a1));
  asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(b0), "f"(b1));
  asm("add.rn.f32x2 %0, %0, %1;" : "+l"(ra) : "l"(rb));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a0), "=f"(a1) : "l"(ra));
}
__device__ __forceinline__ void sub2(float& a0, float& a1, float b0, float b1) {
  unsigned long long ra, rb;
  asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(a0), "f"(a1));
  asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(b0), "f"(b1));
  asm("sub.rn.f32x2 %0, %0, %1;" : "+l"(ra) : "l"(rb));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a0), "=f"(a1) : "l"(ra));
}

struct Result {
  unsigned long long cycles;
  float sink;
  int bad;
};

// offs[] holds warp-uniform "tap offsets" (read from shared memory like the real kernel does).
template <int N, int MODE, int NW>
__global__ void __launch_bounds__(NW * 32, 1) bench_kernel(int iters, Result* out) {
  extern __shared__ __align__(16) float sm[];  // 8192 floats of data + 64 offsets
  __shared__ uint32_t tm_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < 8192; i += blockDim.x) sm[i] = (float)(i & 1023) * 0.001f;
  int* offs = reinterpret_cast<int*>(sm + 8192);
  if (tid < 64) offs[tid] = (tid * 37) % (512 - N);
  if (warp == 0) tmem_alloc512(&tm_slot);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tbase = tm_slot + ((uint32_t)((warp & 3) * 32) << 16);

  // fill this warp's lanes (only the first warp of each lane quarter), row l = l*1000 + column
  int bad = 0;
  if (warp < 4) {
    for (int c0 = 0; c0 < 512; c0 += 32) {
      uint32_t v[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = __float_as_uint((float)((warp * 32 + lane) * 1000 + c0 + j));
      sttm32(tbase + c0, v);
    }
    tmem_wait_st();
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  {  // read back from an unaligned column offset
    uint32_t v[N];
    const int c0 = 13 + warp;
    ldtm<N>(v, tbase + c0);
    tmem_wait_ld();
#pragma unroll
    for (int j = 0; j < N; ++j)
      if (__uint_as_float(v[j]) != (float)(((warp & 3) * 32 + lane) * 1000 + c0 + j)) ++bad;
  }
  __syncthreads();

  float acc[N];
#pragma unroll
  for (int j = 0; j < N; ++j) acc[j] = 0.0f;
  const float* px = sm + lane * 33 + warp * 64;  // odd lane stride: conflict-free scalar LDS
  unsigned long long t0 = clock64();
  if constexpr (MODE == LDTM) {
    for (int it = 0; it < iters; ++it) {
      const int off = offs[it & 63];
      uint32_t v[N];
      ldtm<N>(v, tbase + off);
      tmem_wait_ld();
#pragma unroll
      for (int j = 0; j < N; ++j) acc[j] = __fadd_rn(acc[j], __uint_as_float(v[j]));
    }
  } else if constexpr (MODE == LDTM_NOWAIT2) {  // two loads in flight before the wait
    for (int it = 0; it < iters; it += 2) {
      const int off = offs[it & 63], off2 = offs[(it + 1) & 63];
      uint32_t v[N], u[N];
      ldtm<N>(v, tbase + off);
      ldtm<N>(u, tbase + off2);
      tmem_wait_ld();
#pragma unroll
      for (int j = 0; j < N; ++j) acc[j] = __fadd_rn(acc[j], __uint_as_float(v[j]));
#pragma unroll
      for (int j = 0; j < N; ++j) acc[j] = __fadd_rn(acc[j], __uint_as_float(u[j]));
    }
  } else if constexpr (MODE == LDS) {
    for (int it = 0; it < iters; ++it) {
      const int off = offs[it & 63];
      float t[N];
#pragma unroll
      for (int j = 0; j < N; ++j) t[j] = px[off + j];
#pragma unroll
      for (int j = 0; j < N; ++j) acc[j] = __fadd_rn(acc[j], t[j]);
    }
  } else if constexpr (MODE == MIXED) {
    for (int it = 0; it < iters; it += 2) {
      const int off = offs[it & 63], off2 = offs[(it + 1) & 63];
      uint32_t v[N];
      ldtm<N>(v, tbase + off);
      float t[N];
#pragma unroll
      for (int j = 0; j < N; ++j) t[j] = px[off2 + j];
#pragma unroll
      for (int j = 0; j < N; ++j) acc[j] = __fadd_rn(acc[j], t[j]);
      tmem_wait_ld();
#pragma unroll
      for (int j = 0; j < N; ++j) acc[j] = __fadd_rn(acc[j], __uint_as_float(v[j]));
    }
  } else if constexpr (MODE == FADD1) {  // pure FADD issue rate: N independent chains
    float b = (float)lane;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int j = 0; j < N; ++j) acc[j] = __fadd_rn(acc[j], b);
      b = b + 1.0f;
    }
  } else if constexpr (MODE == FADD2) {  // the same adds as N/2 packed instructions
    float b = (float)lane;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int j = 0; j < N; j += 2) {
        if (it & 1) sub2(acc[j], acc[j + 1], b, b);
        else add2(acc[j], acc[j + 1], b, b);
      }
      b = b + 1.0f;
    }
  } else if constexpr (MODE == LDTM_FADD2) {
    for (int it = 0; it < iters; it += 2) {
      const int off = offs[it & 63], off2 = offs[(it + 1) & 63];
      uint32_t v[N], u[N];
      ldtm<N>(v, tbase + off);
      ldtm<N>(u, tbase + off2);
      tmem_wait_ld();
#pragma unroll
      for (int j = 0; j < N; j += 2) add2(acc[j], acc[j + 1], __uint_as_float(v[j]), __uint_as_float(v[j + 1]));
#pragma unroll
      for (int j = 0; j < N; j += 2) sub2(acc[j], acc[j + 1], __uint_as_float(u[j]), __uint_as_float(u[j + 1]));
    }
  } else if constexpr (MODE == STTM) {
    if constexpr (N == 32) {
      uint32_t v[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = lane + j;
      for (int it = 0; it < iters; ++it) {
        const int off = offs[it & 63];
        sttm32(tbase + off, v);
      }
      tmem_wait_st();
    }
  }
  __syncthreads();
  unsigned long long t1 = clock64();
  float s = 0.0f;
#pragma unroll
  for (int j = 0; j < N; ++j) s += acc[j];
  if (tid == 0) {
    out[blockIdx.x].cycles = t1 - t0;
  }
  if (bad) atomicAdd(&out[blockIdx.x].bad, bad);
  if (s == 123.456f) out[blockIdx.x].sink = s;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) tmem_dealloc512(tm_slot);
}

template <int N, int MODE, int NW>
void run(const char* name, int iters, Result* d_out, int sms) {
  const int nw = NW;
  CK(cudaMemset(d_out, 0, sizeof(Result) * sms));
  size_t smem = (8192 + 64) * 4;
  auto k = bench_kernel<N, MODE, NW>;
  CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  k<<<sms, nw * 32, smem>>>(iters, d_out);  // warm-up
  CK(cudaDeviceSynchronize());
  CK(cudaMemset(d_out, 0, sizeof(Result) * sms));
  CK(cudaEventRecord(e0));
  k<<<sms, nw * 32, smem>>>(iters, d_out);
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  Result* h = (Result*)malloc(sizeof(Result) * sms);
  CK(cudaMemcpy(h, d_out, sizeof(Result) * sms, cudaMemcpyDeviceToHost));
  double cyc = 0;
  int bad = 0;
  for (int i = 0; i < sms; ++i) {
    cyc += (double)h[i].cycles;
    bad += h[i].bad;
  }
  cyc /= sms;
  const double words = (double)nw * 32.0 * iters * N;  // words moved to (from) registers per SM
  printf("%-7s N=%2d NW=%2d iters=%d  cycles=%.0f  %.1f B/clk/SM  %.3f words/clk/SM  (%.3f ms)  readback_bad=%d\n", name, N, nw, iters,
         cyc, words * 4.0 / cyc, words / cyc, ms, bad);
  free(h);
}

template <int NW>
void all(int iters, Result* d_out, int sms) {
  run<32, LDS, NW>("lds", iters, d_out, sms);
  run<16, LDTM, NW>("ldtm", iters, d_out, sms);
  run<32, LDTM, NW>("ldtm", iters, d_out, sms);
  if constexpr (NW <= 8) run<64, LDTM, NW>("ldtm", iters, d_out, sms);
  run<16, LDTM_NOWAIT2, NW>("ldtm2", iters, d_out, sms);
  if constexpr (NW <= 16) run<32, LDTM_NOWAIT2, NW>("ldtm2", iters, d_out, sms);
  if constexpr (NW <= 16) run<32, MIXED, NW>("mixed", iters, d_out, sms);
  else run<16, MIXED, NW>("mixed", iters, d_out, sms);
  run<32, STTM, NW>("sttm", iters, d_out, sms);
  if constexpr (NW <= 16) {
    run<32, FADD1, NW>("fadd", iters, d_out, sms);
    run<32, FADD2, NW>("fadd2", iters, d_out, sms);
    run<16, LDTM_FADD2, NW>("ldtm+f2", iters, d_out, sms);
    run<32, LDTM_FADD2, NW>("ldtm+f2", iters, d_out, sms);
  }
}

int main() {
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  printf("device %s, %d SMs\n", prop.name, sms);
  Result* d_out;
  CK(cudaMalloc(&d_out, sizeof(Result) * sms));
  const int iters = 4096;
  all<4>(iters, d_out, sms);
  all<8>(iters, d_out, sms);
  all<16>(iters, d_out, sms);
  all<32>(iters, d_out, sms);
  return 0;
}
