// Minimal stand-ins for <stdint.h> / <float.h> when the device sources are compiled by NVRTC (no host headers there):
// spec.cpp compiles interp.cu again, specialised for one tape, at bank-build time.
#pragma once
#if defined(__CUDACC_RTC__)
typedef unsigned char uint8_t;
typedef unsigned short uint16_t;
typedef unsigned int uint32_t;
typedef unsigned long long uint64_t;
typedef signed char int8_t;
typedef short int16_t;
typedef int int32_t;
typedef long long int64_t;
#ifndef FLT_MAX
#define FLT_MAX 3.402823466e+38F
#endif
#endif
