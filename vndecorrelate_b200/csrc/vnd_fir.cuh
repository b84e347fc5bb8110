// Parameter block shared by the sparse-FIR kernels.
#pragma once

#include "vnd_common.cuh"

namespace vnd {

enum { MODE_SEG = 0, MODE_ASC32 = 1, MODE_ASC64 = 2 };

struct FirParams {
  const void* x;
  long long x_st, x_sc;
  float* y;
  long long y_st, y_sc;
  long long frames;
  int channels;
  const int* words;
  const int* offsets;
  int tile;
  int halo;
  int tiles_per_channel;
  int apply_gain;
  int bulk_ok;
};


// vnd_fir_window.cu: the register-window variant for planar float32 SEGMENTED programs.  Returns
// VND_EUNSUPPORTED (no error text) when the request does not qualify.
int fir_window_launch(const FirParams& p, int max_prog_words, cudaStream_t st);

// vnd_fir_tmem.cu: the tensor-memory variant for planar float32 SEGMENTED programs.  Covers the
// interior tiles of every channel and reports the frames done per channel; the caller finishes the
// tail.  Returns VND_EUNSUPPORTED (no error text) when the request does not qualify.
int fir_tmem_launch(const FirParams& p, int max_prog_words, cudaStream_t st, long long* frames_done);

// vnd_fir_ring.cu: the ring-buffer variant for long filters on planar float32 SEGMENTED programs (persistent CTAs,
// every sample fetched once under the taps of the previous step).  Covers the interior of every channel and reports
// the frames done per channel; the caller finishes the tail.  VND_EUNSUPPORTED (no error text) when it does not qualify.
int fir_ring_launch(const FirParams& p, int max_prog_words, cudaStream_t st, long long* frames_done);

}  // namespace vnd
