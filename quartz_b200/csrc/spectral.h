// K5 — frame-parallel evaluation of spectral patches (BASELINE configs[3]: the spectral-gate / spectral-delay chain of
// /root/reference/src/nodes.rs:601-700 `Rfft` / `Ifft`).
//
// The reference streams one bin per sample through `rfft -> stateless bin ops -> ifft`; a frame of N bins therefore
// depends only on N input samples, and when everything that feeds the rfft is a pure function of time (counter-based noise,
// wave tables, delays of those) every frame of every instance of every voice can be computed independently.  plan_spectral()
// recognises such tapes and splits them into three mini-tapes per segment; spectral.cu evaluates whole frames in shared
// memory (input samples -> FFT -> bin chain on all N bins -> inverse FFT) and hands the resynthesised signal to the
// post-graph through an L2-resident ring.
#pragma once
#include <cuda_runtime.h>

#include <vector>

#include "lower.h"

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

struct SpPlan {
  bool ok = false;
  std::vector<Instr> code;      // all mini-tapes; source instructions carry their time offset in `pad` (int32)
  std::vector<SpSegment> segs;
  std::vector<SpItem> items;    // the frames of one round of C samples
  int post_lo = 0, post_hi = 0; // mini-tape of the post-graph (reads the Y ring through OP_STREAM_IN)
  std::vector<uint16_t> out_x;  // X index of each net output in the post mini-tape's numbering
  int n_slots_frame = 0;        // temporaries the pre / chain mini-tapes need (per sample of a block)
  int n_slots_post = 0;
  int n_streams = 0;
  int C = 0;                    // round length = the largest transform size
};

SpPlan plan_spectral(const Tape& t);

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
size_t spectral_y_bytes(const SpPlan& p, long V, int* ring);
cudaError_t launch_spectral(const SpArgs& a, const SpPlan& p, cudaStream_t stream, int* launches);

}  // namespace qg
