// Hand-fused kernels (see fused.h).
#include "fused.h"

namespace qg {

FusedPlan plan_fused(const Tape&) { return FusedPlan(); }
const char* fused_name(int id) { return id == FUSED_NOISE_SVF ? "k_noise_svf_scan" : id == FUSED_SINE_SVF_ENV ? "k_polysynth" : "none"; }
cudaError_t launch_fused(const FusedPlan&, const FusedArgs&, cudaStream_t, int*) { return cudaErrorNotSupported; }

}  // namespace qg
