// Shared-memory transforms of the frame-parallel spectral path (K5): included by spectral.cu and by the NVRTC-specialised
// kernels (spectral_kernel.cuh), inside namespace qg.
#pragma once
// Transform buffers hold interleaved complex values (one LDS.64 / STS.64 per point), padded by one element per 16.
#define CPAD(i) ((i) + ((i) >> 4))

// Radix-2 DIT butterflies, 2^R points per thread carried through R consecutive stages in registers (the schedule of
// k_interp_tv's transforms: bit-identical to the plain radix-2 loop of the oracle).  sgn = +1 forward, -1 inverse: the
// inverse conjugates the twiddle, and -(w.y) == w.y * -1 exactly.  LG and S are compile-time: every index below is an
// immediate offset from one per-thread base.  The buffer is addressed as an OFFSET (in float2) into the kernel's shared
// memory, so every access is an LDS / STS whatever the inliner does; TWS: the twiddle table has been staged in shared memory
// at offset tw_off (the specialised kernels do that once per frame), else it is read through the read-only cache.
#define QG_SMEM_C (reinterpret_cast<float2*>(qg_smem))
template <int LG, int S, int R, bool TWS>
__device__ __forceinline__ void sp_fft_pass(int f_off, const float2* __restrict__ tw, int tw_off, float sgn, int tid, int nth) {
  constexpr uint32_t N = 1u << LG, h = 1u << S;
  float2* f = QG_SMEM_C + f_off;
  for (uint32_t g = tid; g < (N >> R); g += nth) {
    const uint32_t k = g & (h - 1), base = ((g >> S) << (S + R)) | k;
    float2 x[1 << R];
#pragma unroll
    for (int m = 0; m < (1 << R); m++) x[m] = f[CPAD(base + (uint32_t)m * h)];
#pragma unroll
    for (int q = 0; q < R; q++) {
      const int hq = 1 << q;
      float2 w[1 << (R - 1)];                      // the 2^q distinct twiddles of stage S + q
#pragma unroll
      for (int e = 0; e < hq; e++) {
        const uint32_t ti = (k + (uint32_t)e * h) << (LG - 1 - (S + q));
        w[e] = TWS ? QG_SMEM_C[tw_off + ti] : __ldg(tw + ti);
        w[e].y *= sgn;
      }
#pragma unroll
      for (int m = 0; m < (1 << R); m++) {
        if (m & hq) continue;
        const float wr = w[m & (hq - 1)].x, wi = w[m & (hq - 1)].y;
        const float2 u = x[m], v = x[m + hq];
        const float tr = v.x * wr - v.y * wi, tim = v.x * wi + v.y * wr;
        x[m] = make_float2(u.x + tr, u.y + tim);
        x[m + hq] = make_float2(u.x - tr, u.y - tim);
      }
    }
#pragma unroll
    for (int m = 0; m < (1 << R); m++) f[CPAD(base + (uint32_t)m * h)] = x[m];
  }
}
template <int LG, int S, bool TWS>
__device__ __forceinline__ void sp_fft_from(int f_off, const float2* tw, int tw_off, float sgn, int tid, int nth) {
  if constexpr (S < LG) {
    constexpr int R = LG - S >= 3 ? 3 : LG - S;
    sp_fft_pass<LG, S, R, TWS>(f_off, tw, tw_off, sgn, tid, nth);
    __syncthreads();
    sp_fft_from<LG, S + R, TWS>(f_off, tw, tw_off, sgn, tid, nth);
  }
}
template <int LG, bool TWS>
__device__ __noinline__ void sp_fft_n(int f_off, const float2* tw, int tw_off, float sgn, int tid, int nth) {
  sp_fft_from<LG, 0, TWS>(f_off, tw, tw_off, sgn, tid, nth);
}
__device__ __forceinline__ void sp_fft(int f_off, int lg, const float2* tw, float sgn, int tid, int nth) {
  switch (lg) {
    case 3: sp_fft_n<3, false>(f_off, tw, 0, sgn, tid, nth); break;
    case 4: sp_fft_n<4, false>(f_off, tw, 0, sgn, tid, nth); break;
    case 5: sp_fft_n<5, false>(f_off, tw, 0, sgn, tid, nth); break;
    case 6: sp_fft_n<6, false>(f_off, tw, 0, sgn, tid, nth); break;
    case 7: sp_fft_n<7, false>(f_off, tw, 0, sgn, tid, nth); break;
    case 8: sp_fft_n<8, false>(f_off, tw, 0, sgn, tid, nth); break;
    case 9: sp_fft_n<9, false>(f_off, tw, 0, sgn, tid, nth); break;
    case 10: sp_fft_n<10, false>(f_off, tw, 0, sgn, tid, nth); break;
    case 11: sp_fft_n<11, false>(f_off, tw, 0, sgn, tid, nth); break;
    default: sp_fft_n<12, false>(f_off, tw, 0, sgn, tid, nth); break;
  }
}
