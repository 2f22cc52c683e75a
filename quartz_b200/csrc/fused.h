// Hand-fused kernels for hot tape shapes (K2 time-parallel noise->filter banks, the poly-synth voice, ...).
// plan_fused() pattern-matches a lowered tape; when it matches, launch_fused() runs the specialised kernel
// with all per-voice state in registers instead of the generic interpreter.
#pragma once
#include <cuda_runtime.h>

#include "lower.h"

namespace qg {

enum FusedId : int { FUSED_NONE = 0, FUSED_NOISE_SVF = 1, FUSED_SINE_SVF_ENV = 2 };

struct FusedPlan {
  int id = FUSED_NONE;
  int p[16];   // parameter indices the kernel reads (meaning depends on id)
  int s[16];   // state indices (relative to the state region)
};
struct FusedArgs {
  const float* params;   // [P][Vp]
  float* state;          // [NS][Vp]
  int V, Vp;
  long T;
  int group;
  float* out;
  float** scratch;        // bank-owned scratch buffer (segment states), grown on demand
  size_t* scratch_bytes;
  float sample_rate;
  const float* tables;    // the bank's table region (wavetable sets)
};

FusedPlan plan_fused(const Tape& t);
const char* fused_name(int id);
cudaError_t launch_fused(const FusedPlan& plan, const FusedArgs& a, cudaStream_t stream, int* launches);
// FUSED_SINE_SVF_ENV with the sine oscillator: packed two-voices-per-lane kernel (fused_poly.cu), 64-sample envelope windows
cudaError_t launch_polysynth_x2(const FusedPlan& plan, const FusedArgs& a, cudaStream_t stream);

}  // namespace qg
