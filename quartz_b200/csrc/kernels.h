// Kernel argument blocks and launchers (device code lives in interp.cu / fused.cu / fft.cu).
#pragma once
#if defined(__CUDACC_RTC__)
#include "rtc_compat.h"
#else
#include <cuda_runtime.h>
#include <stdint.h>
#endif

#include "tape.h"

namespace qg {

struct InterpArgs {
  const Instr* code;
  int n_instr;
  int P, NS, NT;              // parameter / state / temporary counts (X = [P | NS | NT])
  int n_in, n_out;
  const uint16_t* out_x;      // X index of each output
  const float* params;        // [P][Vp]
  float* state;               // [NS][Vp]
  const float* state_init;    // [NS][Vp]
  const uint8_t* state_keep;  // [NS] 1: word survives a reset of the nested net it belongs to
  float* rings;               // [ring_floats][Vp]
  const Ring* ring_tab;
  const ResetRange* resets;
  const float* tables;
  const float* in;            // net inputs (process/apply): voice-major [V][n_in][T] or frame-major [T][V][n_in]
  float* out;                 // voice-major [V/group][n_out][T] or frame-major [T][V][n_out]
  int V, Vp;                  // voices, voices padded to a multiple of 128
  long T;
  int in_frame_major, out_frame_major;
  int group;                  // K6: sum consecutive voices in groups of `group` (1, 2, 4, 8, 16 or 32), scaled by 1/group
};

// time-vector mode (one CTA per voice, threads = samples of a hop); rings are voice-major [V][ring_floats]
struct TvArgs {
  const Instr* code;
  int n_instr;
  int P, NS, NT;
  int n_in, n_out;
  const uint16_t* out_x;
  const float* params;        // [P][Vp]
  float* state;               // [NS][Vp]
  float* rings;               // [V][ring_floats]
  uint32_t ring_floats;
  const Ring* ring_tab;
  const float* tables;
  const float* in;            // voice-major [V][n_in][T]
  float* out;                 // voice-major [V][n_out][T]
  int V, Vp;
  long T;
  int H;                      // hop (samples per pass), divides every FFT size and start offset
  int fft_n;                  // largest FFT size in the tape (shared-memory transform buffer), 0 if none
  int frame_major;            // 1: in/out are frame-major [T][V][ch] instead of voice-major [V][ch][T]
  int align_s;                // X index of one rfft/ifft counter (hop alignment across calls), -1 if none
  int n_lti;                  // fixed-coefficient LTI filters in the tape (scan-matrix table in shared memory)
  int biquad_scan;            // 0: some voice's direct-form biquad is too ill-conditioned to re-associate -> one-thread exact order
  // shared-memory layout in floats, filled by launch_interp_tv(): kernel parameters live in the constant bank and fold into
  // the consuming instructions, whereas values derived in the kernel were being re-materialised at every use
  int PS, ps_off, tmp_off, oldv_off, fr_off, fi_off, lti_off, scan_off, segi_off, segt_off;
};
#if !defined(__CUDACC_RTC__)   // launchers: host side only
size_t tv_smem_bytes(const TvArgs& a);
cudaError_t launch_interp_tv(const TvArgs& a, cudaStream_t stream, int* launches);

// block_ok: the tape may run on the block-mode lane interpreter (feed-forward, every feedback ring >= its block length)
cudaError_t launch_interp(const InterpArgs& a, bool divergent, bool block_ok, cudaStream_t stream, int* launches);
int interp_block_len();
// 64 flop x iters per thread, blocks x 256 threads
cudaError_t launch_fp32_peak(float* out, int blocks, int iters, cudaStream_t stream);
cudaError_t launch_init_state(float* state_init, const uint32_t* defaults, int NS, int Vp, const HashInit* hi, int n_hi,
                              const uint64_t* salts, cudaStream_t stream);
// AudioUnit::reset for every voice: state <- state_init except the words marked in `keep`
cudaError_t launch_reset_state(float* state, const float* state_init, const uint8_t* keep, int NS, int Vp, cudaStream_t stream);
cudaError_t launch_broadcast_params(float* params, const float* tmpl, int P, int Vp, cudaStream_t stream);
cudaError_t launch_stereo_frames(const float* src, int n_ch, long n, float* frames, cudaStream_t stream);
// i16 (offset_binary = 0) or u16 (1) frames, interleaved L R
cudaError_t launch_stereo_frames_i16(const float* src, int n_ch, long n, int offset_binary, void* frames, cudaStream_t stream);
cudaError_t launch_mix_rows(const float* rows, int R, long T, float scale, float* out, cudaStream_t stream);
#endif

}  // namespace qg
