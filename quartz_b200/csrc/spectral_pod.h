// Plain-data structures of the frame-parallel spectral path (K5), shared by the host planner, the offline-compiled kernels
// (spectral.cu) and the NVRTC-specialised ones (spectral_kernel.cuh).
#pragma once
#if defined(__CUDACC_RTC__)
#include "rtc_compat.h"
#else
#include <stdint.h>
#endif

#include "tape.h"

namespace qg {

struct SpSegment {      // one `rfft(N, start) -> chain -> ifft(N, start)` instance
  int lg;               // N = 1 << lg
  int start;            // initial counter of both nodes (nodes.rs:609-616, 668-675)
  int pre_lo, pre_hi;   // mini-tape range that evaluates the rfft input at time t' (random access in time)
  int pre_x;            // X index of that value (a scalar when the input is a constant)
  int ch_lo, ch_hi;     // the bin chain, operands renumbered into a compact slot space
  int rf_x;             // X index of the chain's inputs (re, im = rf_x, rf_x + 1)
  int in_re_x, in_im_x; // X indices of the ifft's two inputs (chain outputs, scalars or the rfft outputs themselves)
  int tw;               // twiddle table offset in the bank's table region
  int y_re, y_im;       // stream index of the ifft's outputs in the Y ring, -1 when the post-graph never reads it
  int sym_re, sym_im;   // the chain's results under conjugation of the input bin: +1 unchanged, -1 negated; 0 0 = unknown
};
struct SpItem { int seg, frame; };   // work item of one round: frame `frame` (0 .. C/N - 1) of segment `seg`

struct SpArgs {
  const Instr* code; int n_code;
  const SpSegment* segs; int n_segs;
  const SpItem* items; int n_items;
  const uint16_t* out_x; int n_out;
  const float* params;        // [P][Vp]
  const float* state_init;    // [NS][Vp]: K5 evaluates from the state at reset plus an absolute sample time
  const float* tables;
  int P, NS, V, Vp;
  float* y;                   // [n_streams][V][ring]
  int ring;                   // samples per Y row, a power of two >= 2 C
  int n_streams;
  float* out;
  long T;                     // samples of this call
  long t0;                    // absolute time (samples since reset) of the call's first sample
  int frame_major;
  int C, post_lo, post_hi, n_slots_frame, n_slots_post;
};

}  // namespace qg
