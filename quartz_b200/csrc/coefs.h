// Coefficient formulas shared by the host lowering (fixed-parameter nodes: computed once per voice with the
// host libm, like the reference would) and the device (variable-input nodes: recomputed when an input moves).
// FunDSP 0.18.2 algorithms restated from the published designs (Simper SVF, RBJ/Butterworth biquads, one-pole
// filters); the crate source is not vendored in /root/reference — see DESIGN.md "parity tiers".
#pragma once
#if !defined(__CUDACC_RTC__)
#include <math.h>
#endif

#if defined(__CUDACC__)
#define QG_HD __host__ __device__ __forceinline__
#else
#define QG_HD inline
#endif

namespace qg {

static const float QG_PI = 3.14159265358979323846f;
static const float QG_TAU = 6.28318530717958647692f;

// mode: 0 lowpass 1 highpass 2 bandpass 3 notch 4 peak 5 allpass 6 bell 7 lowshelf 8 highshelf
// c[0..5] = a1 a2 a3 m0 m1 m2
QG_HD void svf_coefs(int mode, float cutoff, float q, float gain, float sr, float* c) {
  float g, k, A;
  switch (mode) {
    case 6: A = sqrtf(gain); g = tanf(QG_PI * cutoff / sr); k = 1.0f / (q * A); break;
    case 7: A = sqrtf(gain); g = tanf(QG_PI * cutoff / sr) / sqrtf(A); k = 1.0f / q; break;
    case 8: A = sqrtf(gain); g = tanf(QG_PI * cutoff / sr) * sqrtf(A); k = 1.0f / q; break;
    default: A = 1.0f; g = tanf(QG_PI * cutoff / sr); k = 1.0f / q; break;
  }
  float a1 = 1.0f / (1.0f + g * (g + k));
  float a2 = g * a1;
  float a3 = g * a2;
  float m0, m1, m2;
  switch (mode) {
    case 0: m0 = 0; m1 = 0; m2 = 1; break;
    case 1: m0 = 1; m1 = -k; m2 = -1; break;
    case 2: m0 = 0; m1 = 1; m2 = 0; break;
    case 3: m0 = 1; m1 = -k; m2 = 0; break;
    case 4: m0 = 1; m1 = -k; m2 = -2; break;
    case 5: m0 = 1; m1 = -2.0f * k; m2 = 0; break;
    case 6: m0 = 1; m1 = k * (A * A - 1.0f); m2 = 0; break;
    case 7: m0 = 1; m1 = k * (A - 1.0f); m2 = A * A - 1.0f; break;
    default: m0 = A * A; m1 = k * (1.0f - A) * A; m2 = 1.0f - A * A; break;
  }
  c[0] = a1; c[1] = a2; c[2] = a3; c[3] = m0; c[4] = m1; c[5] = m2;
}

// kind 1 butterpass(hz), 2 resonator(hz, bandwidth); c[0..4] = a1 a2 b0 b1 b2
QG_HD void biquad_coefs(int kind, float p0, float p1, float sr, float* c) {
  if (kind == 1) {
    float f = tanf(p0 * QG_PI / sr);
    float a0r = 1.0f / (1.0f + 1.41421356237309504880f * f + f * f);
    c[0] = (2.0f * f * f - 2.0f) * a0r;
    c[1] = (1.0f - 1.41421356237309504880f * f + f * f) * a0r;
    c[2] = f * f * a0r;
    c[3] = 2.0f * c[2];
    c[4] = c[2];
  } else {
    float r = expf(-QG_PI * p1 / sr);
    c[0] = -2.0f * r * cosf(QG_TAU * p0 / sr);
    c[1] = r * r;
    c[2] = sqrtf(1.0f - r * r) * 0.5f;
    c[3] = 0.0f;
    c[4] = -c[2];
  }
}

// kind 0 lowpole 1 highpole 2 dcblock 3 allpole(delay in samples)
QG_HD float onepole_coef(int kind, float p, float sr) {
  if (kind == 0 || kind == 1) return expf(-QG_TAU * p / sr);
  if (kind == 2) return 1.0f - QG_TAU * p / sr;
  return (1.0f - p) / (1.0f + p);
}

// equal-power pan weights
QG_HD void pan_weights(float pan, float* l, float* r) {
  float p = pan;
  if (p < -1.0f) p = -1.0f;
  if (p > 1.0f) p = 1.0f;
  float a = (p + 1.0f) * (QG_PI * 0.25f);
  *l = cosf(a);
  *r = sinf(a);
}

}  // namespace qg
