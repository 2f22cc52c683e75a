// Tape specialisation (K1s): a lane kernel compiled for ONE tape with NVRTC, see spec_kernel.cuh.
#pragma once
#include <string>

#include "lower.h"

namespace qg {

// Uniform tapes without per-lane FFT nodes can be specialised; `why` names the first obstacle otherwise.
bool spec_supported(const Tape& t, std::string* why);
// The translation unit NVRTC compiles for this tape (includes interp.cu and spec_kernel.cuh by name: the include path is
// the csrc/ directory next to the library).
std::string spec_source(const Tape& t);

}  // namespace qg

#if !defined(QG_SPEC_HOST_ONLY)
#include <cuda_runtime.h>

#include "kernels.h"
#include "spectral.h"

namespace qg {

struct SpecKernel {
  cudaLibrary_t lib = nullptr;
  cudaKernel_t fn = nullptr;
  double compile_seconds = 0.0;
  bool shared = false;   // owned by the process-wide cache (spec.cpp): never unloaded by a bank
};
// Compiles spec_source(t) for sm_100a with the NVRTC found at run time (dlopen: the library has no link-time dependency on
// it) and loads the cubin.  Returns false with `err` set when NVRTC is missing, the sources next to the library are not
// found, or compilation fails — the caller keeps the interpreter.
bool spec_compile(const Tape& t, SpecKernel* out, std::string* err);
// AUTO policy helpers: the tape's specialised kernel would not pay a dependent HBM load per sample; the kernel is already in
// the process-wide cache (no compile needed)
bool spec_auto_ok(const Tape& t);
bool spec_cached(const Tape& t, SpecKernel* out);
cudaError_t spec_launch(const SpecKernel& k, const InterpArgs& a, cudaStream_t stream, int* launches);
void spec_release(SpecKernel* k);

// K5s: the frame-parallel spectral kernels compiled for one plan (spectral_kernel.cuh); cached process-wide by source, never
// unloaded.  Same failure modes as spec_compile: the caller keeps the generic kernels of spectral.cu.
struct SpectralKernels {
  cudaLibrary_t lib = nullptr;
  cudaKernel_t frames = nullptr, post = nullptr;
  double compile_seconds = 0.0;
};
std::string spectral_spec_source(const SpPlan& p, const Tape& t);
bool spectral_spec_compile(const SpPlan& p, const Tape& t, SpectralKernels* out, std::string* err);
bool spectral_spec_cached(const SpPlan& p, const Tape& t, SpectralKernels* out);   // already compiled in this process

}  // namespace qg
#endif
