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
#include "spectral_pod.h"

namespace qg {

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

size_t spectral_y_bytes(const SpPlan& p, long V, int* ring);
// spec_frames / spec_post: the plan's NVRTC-specialised kernels (spectral_kernel.cuh) or null for the generic ones
cudaError_t launch_spectral(const SpArgs& a, const SpPlan& p, cudaStream_t stream, int* launches, cudaKernel_t spec_frames = nullptr,
                            cudaKernel_t spec_post = nullptr);

}  // namespace qg
