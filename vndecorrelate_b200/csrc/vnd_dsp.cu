// Analysis-side helpers of the reference's utils/dsp.py on the device (sm_100a): the normalisers in every mode and the
// polar form of a stereo signal.  None of this is on the throughput path (VelvetNoise uses rms_normalize in DUAL_MONO
// mode on 2-D signals, which vnd_post.cu serves with the numpy axis-0 ORDER kernels); these entry points complete the
// drop-in surface:
//   rms_normalize   src/vndecorrelate/utils/dsp.py:87-109   STEREO mode and 1-D signals reduce with axis=None
//   peak_normalize  src/vndecorrelate/utils/dsp.py:71-84
//   polar_coordinates src/vndecorrelate/utils/dsp.py:374-422
//
// What "the same result" needs here: a reduction with axis=None over a contiguous array is numpy's PAIRWISE sum
// (numpy/_core/src/umath/loops_utils.h.src, *_pairwise_sum): blocks of at most 128 elements are summed in eight
// interleaved accumulators that are combined as ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)) plus a sequential remainder, and
// longer ranges are split at n/2 rounded down to a multiple of 8, recursively.  The tree depends on n only, so it is
// evaluated here in parallel and bit for bit: one thread per leaf block, then the fixed combine tree.  Maxima are
// order-independent; everything else is elementwise IEEE arithmetic in numpy's dtypes.

#include "vnd_common.cuh"

namespace vnd {

namespace {

constexpr int PW_BLOCK = 128;  // numpy's PW_BLOCKSIZE

template <typename T, bool SQUARE>
__device__ __forceinline__ T pw_elem(const T* __restrict__ a, long long i) {
  const T v = a[i];
  if constexpr (SQUARE) {
    if constexpr (sizeof(T) == 8) return dmul(v, v);
    else return fmul(v, v);
  } else {
    return v;
  }
}
template <typename T>
__device__ __forceinline__ T pw_add(T a, T b) {
  if constexpr (sizeof(T) == 8) return dadd(a, b);
  else return fadd(a, b);
}

// numpy's leaf: n <= 128 elements starting at a[off]
template <typename T, bool SQUARE>
__device__ T pw_leaf(const T* __restrict__ a, long long off, int n) {
  if (n < 8) {
    T res = (T)0;  // numpy starts the short loop at +0 (the reduction's -0.0 identity is added outside)
    for (int i = 0; i < n; ++i) res = pw_add(res, pw_elem<T, SQUARE>(a, off + i));
    return res;
  }
  T r[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) r[j] = pw_elem<T, SQUARE>(a, off + j);
  int i = 8;
  for (; i < n - (n % 8); i += 8) {
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = pw_add(r[j], pw_elem<T, SQUARE>(a, off + i + j));
  }
  T res = pw_add(pw_add(pw_add(r[0], r[1]), pw_add(r[2], r[3])), pw_add(pw_add(r[4], r[5]), pw_add(r[6], r[7])));
  for (; i < n; ++i) res = pw_add(res, pw_elem<T, SQUARE>(a, off + i));
  return res;
}

// The leaf of numpy's recursion over [0, n) that contains position pos: start and length.
__device__ __forceinline__ void pw_find_leaf(long long n, long long pos, long long* start, int* len) {
  long long off = 0;
  while (n > PW_BLOCK) {
    long long n2 = n / 2;
    n2 -= n2 % 8;
    if (pos < off + n2) {
      n = n2;
    } else {
      off += n2;
      n -= n2;
    }
  }
  *start = off;
  *len = (int)n;
}

// One thread per 64-element window: leaves are 64..128 elements long, so at most one leaf STARTS in a window; the
// thread that owns the window of a leaf's start sums the leaf and parks the result at leafsum[start / 64].
template <typename T, bool SQUARE>
__global__ void pw_leaf_kernel(const T* __restrict__ a, long long n, T* __restrict__ leafsum) {
  const long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long lo = w * 64;
  if (lo >= n) return;
  long long hi = lo + 63;
  if (hi >= n) hi = n - 1;
  long long start;
  int len;
  pw_find_leaf(n, hi, &start, &len);
  if (start >= lo) {
    leafsum[start / 64] = pw_leaf<T, SQUARE>(a, start, len);
  } else if (lo == 0) {  // unreachable: the first leaf starts at 0; kept for clarity
    leafsum[0] = pw_leaf<T, SQUARE>(a, 0, len);
  }
}

// The combine tree, one thread with an explicit stack (depth <= 64): sum(node) = sum(left) + sum(right).
template <typename T>
__global__ void pw_combine_kernel(long long n, const T* __restrict__ leafsum, T* __restrict__ out) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  if (n <= 0) {
    *out = (T)0;
    return;
  }
  struct Frame {
    long long off, n;
    T left;
    int state;  // 0: descend left, 1: left done -> descend right, 2: both done
  };
  Frame st[64];
  int sp = 0;
  st[0] = {0, n, (T)0, 0};
  T ret = (T)0;
  while (sp >= 0) {
    Frame& f = st[sp];
    if (f.n <= PW_BLOCK) {
      ret = leafsum[f.off / 64];
      --sp;
      continue;
    }
    long long n2 = f.n / 2;
    n2 -= n2 % 8;
    if (f.state == 0) {
      f.state = 1;
      st[sp + 1] = {f.off, n2, (T)0, 0};
      ++sp;
    } else if (f.state == 1) {
      f.left = ret;
      f.state = 2;
      st[sp + 1] = {f.off + n2, f.n - n2, (T)0, 0};
      ++sp;
    } else {
      ret = pw_add(f.left, ret);
      --sp;
    }
  }
  *out = ret;
}

// max |a[i * stride]| over i < n (order-independent); one CTA, out[0]
template <typename T>
__global__ void __launch_bounds__(1024) absmax_kernel(const T* __restrict__ a, long long n, long long stride, T* __restrict__ out) {
  __shared__ T red[32];
  T m = (T)0;
  bool nan = false;
  for (long long i = threadIdx.x; i < n; i += blockDim.x) {
    const T v = a[i * stride];
    const T av = v < 0 ? -v : v;
    if (v != v) nan = true;
    if (av > m) m = av;
  }
  for (int o = 16; o > 0; o >>= 1) {
    const T other = __shfl_xor_sync(0xffffffffu, m, o);
    if (other > m) m = other;
    nan = nan || __shfl_xor_sync(0xffffffffu, (int)nan, o);
  }
  __shared__ int any_nan;
  if (threadIdx.x == 0) any_nan = 0;
  __syncthreads();
  if (nan && (threadIdx.x & 31) == 0) any_nan = 1;
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    T r = red[0];
    for (int w = 1; w < (int)((blockDim.x + 31) / 32); ++w)
      if (red[w] > r) r = red[w];
    if (any_nan) r = (T)NAN;  // np.max propagates NaN
    *out = r;
  }
}

// a[i * stride] *= *factor
template <typename T>
__global__ void scale_by_kernel(T* __restrict__ a, long long n, long long stride, const T* __restrict__ factor) {
  const T f = *factor;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    if constexpr (sizeof(T) == 8) a[i * stride] = dmul(a[i * stride], f);
    else a[i * stride] = fmul(a[i * stride], f);
  }
}
template <typename T>
__global__ void divide_by_kernel(T* __restrict__ a, long long n, const T* __restrict__ denom) {
  const T d = *denom;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) a[i] = a[i] / d;
}
template <typename T>
__global__ void divide_into_kernel(const T* __restrict__ a, T* __restrict__ out, long long n, const T* __restrict__ denom) {
  const T d = *denom;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) out[i] = a[i] / d;
}

// Scalar glue, one thread.  op 0: v[0] = v[0] + eps (np.float32 + Python float stays float32)
//                           op 1: v[0] = 1 / (v[0] + eps)                         (peak_normalize)
//                           op 2: mean = fl(sum / T(count))                        (np.mean with axis=None: ret.dtype.type(ret / rcount))
template <typename T>
__global__ void scalar_op_kernel(T* v, int op, double eps, long long count) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const T e = (T)eps;
  if (op == 0) v[0] = pw_add(v[0], e);
  else if (op == 1) v[0] = (T)1 / pw_add(v[0], e);
  else v[0] = v[0] / (T)count;
}

// gains[c] = sqrt(in_stat[c * in_step]) / sqrt(out_stat[c * out_step] + eps), c < n_gain  (utils/dsp.py:107-109)
template <typename T>
__global__ void rms_gain_general_kernel(const T* in_stat, int in_step, const T* out_stat, int out_step, T* gains, int n_gain, double eps) {
  const int c = threadIdx.x;
  if (c >= n_gain || blockIdx.x != 0) return;
  const T num = sizeof(T) == 8 ? (T)sqrt((double)in_stat[c * in_step]) : (T)__fsqrt_rn((float)in_stat[c * in_step]);
  const T den_in = pw_add(out_stat[c * out_step], (T)eps);
  const T den = sizeof(T) == 8 ? (T)sqrt((double)den_in) : (T)__fsqrt_rn((float)den_in);
  gains[c] = num / den;
}

// theta (folded or not), radius (utils/dsp.py:399-413)
template <typename T>
__global__ void polar_kernel(const T* __restrict__ l, const T* __restrict__ r, long long n, int mode_ms, int semicircular, T* __restrict__ radii,
                             T* __restrict__ thetas) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const T a = l[i], b = r[i];
    T th;
    if constexpr (sizeof(T) == 8) {
      th = mode_ms ? atan2(dsub(a, b), dadd(a, b)) : atan2(a, b);
      if (semicircular) {
        const double hp = 1.5707963267948966, pi = 3.141592653589793;  // np.pi / 2, np.pi (float64 comparisons and shifts)
        if (th < -hp) th = dadd(th, pi);
        else if (th > hp) th = dsub(th, pi);
      }
      radii[i] = sqrt(dadd(dmul(a, a), dmul(b, b)));
    } else {
      th = mode_ms ? atan2f(fsub(a, b), fadd(a, b)) : atan2f(a, b);
      if (semicircular) {
        // np.where(thetas < -np.pi / 2, thetas + np.pi, ...): the float32 array meets Python floats: float32 arithmetic
        const float hp = 1.57079637f, pi = 3.14159274f;
        if (th < -hp) th = fadd(th, pi);
        else if (th > hp) th = fsub(th, pi);
      }
      radii[i] = __fsqrt_rn(fadd(fmul(a, a), fmul(b, b)));
    }
    thetas[i] = th;
  }
}

unsigned blocks_for(long long n, int per = 256, long long cap = 148 * 16) {
  long long b = (n + per - 1) / per;
  if (b > cap) b = cap;
  return (unsigned)(b < 1 ? 1 : b);
}

// numpy-order pairwise sum of (squares of) a contiguous range; `leafsum` holds n / 64 + 2 values
template <typename T, bool SQUARE>
int pairwise_launch(const T* a, long long n, T* leafsum, T* out, cudaStream_t st) {
  if (n > 0) {
    const long long windows = (n + 63) / 64;
    pw_leaf_kernel<T, SQUARE><<<blocks_for(windows, 128, 1LL << 30), 128, 0, st>>>(a, n, leafsum);
    int rc = after_launch("pw_leaf_kernel");
    if (rc) return rc;
  }
  pw_combine_kernel<T><<<1, 32, 0, st>>>(n, leafsum, out);
  return after_launch("pw_combine_kernel");
}

}  // namespace

size_t dsp_workspace_bytes(long long elems) { return ((size_t)(elems / 64 + 4) + 64) * 8 * 2 + 1024; }

// numpy-order PAIRWISE sum of squares of one contiguous column (what np.mean(np.square(a), axis=0) does for a column of a
// Fortran-ordered / planar array: numpy puts the inner loop on the axis with the smallest stride).  `leaf` holds n / 64 + 4
// elements of the column's type.
int pairwise_sumsq_launch(const void* col, int dtype, long long n, void* leaf, void* out, cudaStream_t st) {
  if (dtype == VND_F64) return pairwise_launch<double, true>((const double*)col, n, (double*)leaf, (double*)out, st);
  return pairwise_launch<float, true>((const float*)col, n, (float*)leaf, (float*)out, st);
}

// rms_normalize in every shape / mode combination (utils/dsp.py:87-109).  x and y: contiguous C-order (frames, ch) or
// 1-D, same dtype.  axis=None when the signal is 1-D or the mode is STEREO, axis=0 otherwise.
template <typename T>
static int rms_general(const vnd_signal* x, const vnd_signal* y, int x_ndim, int y_ndim, int stereo_mode, double eps, void* ws, cudaStream_t st) {
  const bool x_flat = x_ndim == 1 || stereo_mode, y_flat = y_ndim == 1 || stereo_mode;
  const long long nx = x->frames * (x_ndim == 1 ? 1 : x->channels), ny = y->frames * (y_ndim == 1 ? 1 : y->channels);
  T* scal = reinterpret_cast<T*>(ws);  // [0..1] input stats, [2..3] output stats, [4..5] gains
  T* leaf = scal + 64;
  int rc;
  const int cx = x_flat ? 1 : x->channels, cy = y_flat ? 1 : y->channels;
  VND_REQUIRE(cx <= 2 && cy <= 2, VND_EUNSUPPORTED, "rms_normalize supports mono and stereo signals");
  if (x_flat) {
    if ((rc = pairwise_launch<T, true>(reinterpret_cast<const T*>(x->data), nx, leaf, scal + 0, st))) return rc;
    scalar_op_kernel<T><<<1, 1, 0, st>>>(scal + 0, 2, 0.0, nx);
    if ((rc = after_launch("scalar_op_kernel"))) return rc;
  }
  if (y_flat) {
    if ((rc = pairwise_launch<T, true>(reinterpret_cast<const T*>(y->data), ny, leaf, scal + 2, st))) return rc;
    scalar_op_kernel<T><<<1, 1, 0, st>>>(scal + 2, 2, 0.0, ny);
    if ((rc = after_launch("scalar_op_kernel"))) return rc;
  }
  (void)cx;
  const int n_gain = cy;  // the gain broadcasts against the output
  rms_gain_general_kernel<T><<<1, 32, 0, st>>>(scal + 0, x_flat ? 0 : 1, scal + 2, y_flat ? 0 : 1, scal + 4, n_gain, eps);
  if ((rc = after_launch("rms_gain_general_kernel"))) return rc;
  if (y_flat) {
    scale_by_kernel<T><<<blocks_for(ny), 256, 0, st>>>(reinterpret_cast<T*>(y->data), ny, 1, scal + 4);
    return after_launch("scale_by_kernel");
  }
  for (int c = 0; c < cy; ++c) {
    scale_by_kernel<T><<<blocks_for(y->frames), 256, 0, st>>>(reinterpret_cast<T*>(y->data) + c, y->frames, y->channels, scal + 4 + c);
    if ((rc = after_launch("scale_by_kernel"))) return rc;
  }
  return VND_OK;
}

int rms_normalize_launch(const vnd_signal* x, const vnd_signal* y, int x_ndim, int y_ndim, int stereo_mode, double eps, void* ws, cudaStream_t st) {
  // axis-0 statistics (2-D, DUAL_MONO) use numpy's sequential order: per-column means through the order kernels
  const bool x_flat = x_ndim == 1 || stereo_mode, y_flat = y_ndim == 1 || stereo_mode;
  if (!x_flat || !y_flat) {
    // mixed shapes (one 1-D, one 2-D in DUAL_MONO mode): the axis-0 side comes from the order kernels as mean per column
    return VND_EUNSUPPORTED;
  }
  if (x->dtype == VND_F64) return rms_general<double>(x, y, x_ndim, y_ndim, stereo_mode, eps, ws, st);
  return rms_general<float>(x, y, x_ndim, y_ndim, stereo_mode, eps, ws, st);
}

template <typename T>
static int peak_general(const vnd_signal* y, int ndim, int stereo_mode, double eps, void* ws, cudaStream_t st) {
  T* scal = reinterpret_cast<T*>(ws);
  T* data = reinterpret_cast<T*>(y->data);
  int rc;
  if (ndim == 1 || stereo_mode) {
    const long long n = y->frames * (ndim == 1 ? 1 : y->channels);
    absmax_kernel<T><<<1, 1024, 0, st>>>(data, n, 1, scal);
    if ((rc = after_launch("absmax_kernel"))) return rc;
    scalar_op_kernel<T><<<1, 1, 0, st>>>(scal, 1, eps, 0);
    if ((rc = after_launch("scalar_op_kernel"))) return rc;
    scale_by_kernel<T><<<blocks_for(n), 256, 0, st>>>(data, n, 1, scal);
    return after_launch("scale_by_kernel");
  }
  for (int c = 0; c < y->channels; ++c) {
    absmax_kernel<T><<<1, 1024, 0, st>>>(data + c, y->frames, y->channels, scal + c);
    if ((rc = after_launch("absmax_kernel"))) return rc;
    scalar_op_kernel<T><<<1, 1, 0, st>>>(scal + c, 1, eps, 0);
    if ((rc = after_launch("scalar_op_kernel"))) return rc;
    scale_by_kernel<T><<<blocks_for(y->frames), 256, 0, st>>>(data + c, y->frames, y->channels, scal + c);
    if ((rc = after_launch("scale_by_kernel"))) return rc;
  }
  return VND_OK;
}

int peak_normalize_launch(const vnd_signal* y, int ndim, int stereo_mode, double eps, void* ws, cudaStream_t st) {
  VND_REQUIRE(ndim == 1 || stereo_mode || y->channels <= 32, VND_EUNSUPPORTED, "too many channels");
  if (y->frames * (long long)(ndim == 1 ? 1 : y->channels) == 0) return VND_OK;
  if (y->dtype == VND_F64) return peak_general<double>(y, ndim, stereo_mode, eps, ws, st);
  return peak_general<float>(y, ndim, stereo_mode, eps, ws, st);
}

template <typename T>
static int polar_general(const T* l, const T* r, long long n, int mode_ms, int semicircular, int normalize, T* radii, T* thetas, T* weights,
                         void* ws, cudaStream_t st) {
  T* scal = reinterpret_cast<T*>(ws);
  T* leaf = scal + 64;
  int rc;
  if (n == 0) return VND_OK;
  polar_kernel<T><<<blocks_for(n), 256, 0, st>>>(l, r, n, mode_ms, semicircular, radii, thetas);
  if ((rc = after_launch("polar_kernel"))) return rc;
  if (normalize) {  // radii /= radii.max() + EPSILON
    absmax_kernel<T><<<1, 1024, 0, st>>>(radii, n, 1, scal);
    if ((rc = after_launch("absmax_kernel"))) return rc;
    scalar_op_kernel<T><<<1, 1, 0, st>>>(scal, 0, 1e-10, 0);
    if ((rc = after_launch("scalar_op_kernel"))) return rc;
    divide_by_kernel<T><<<blocks_for(n), 256, 0, st>>>(radii, n, scal);
    if ((rc = after_launch("divide_by_kernel"))) return rc;
  }
  if (weights) {  // weights = radii / (radii.sum() + EPSILON), the sum in numpy's pairwise order
    if ((rc = pairwise_launch<T, false>(radii, n, leaf, scal + 1, st))) return rc;
    scalar_op_kernel<T><<<1, 1, 0, st>>>(scal + 1, 0, 1e-10, 0);
    if ((rc = after_launch("scalar_op_kernel"))) return rc;
    divide_into_kernel<T><<<blocks_for(n), 256, 0, st>>>(radii, weights, n, scal + 1);
    if ((rc = after_launch("divide_into_kernel"))) return rc;
  }
  return VND_OK;
}

int polar_launch(const void* l, const void* r, long long n, int dtype, int mode_ms, int semicircular, int normalize, void* radii, void* thetas,
                 void* weights, void* ws, cudaStream_t st) {
  if (dtype == VND_F64)
    return polar_general<double>((const double*)l, (const double*)r, n, mode_ms, semicircular, normalize, (double*)radii, (double*)thetas,
                                 (double*)weights, ws, st);
  return polar_general<float>((const float*)l, (const float*)r, n, mode_ms, semicircular, normalize, (float*)radii, (float*)thetas, (float*)weights, ws,
                              st);
}

}  // namespace vnd
