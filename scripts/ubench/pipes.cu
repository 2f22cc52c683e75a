// Micro-benchmark of the sm_100a issue/pipe costs the fused kernels are built around (profiles/r02_ubench_pipes.txt):
// for each op, cycles per warp-instruction with W warps per SM sub-partition and C independent chains per thread.
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu ; run: ./pipes
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

#define ITERS 4096
enum Op { FADD_, FFMA_, FADD2_, FFMA2_, FMUL2_, FSETBF, VIMNMX_, FMNMX_, FSETP_SEL, FSETP_PADD, MUFUSIN, MUFUEX2, IMAD_, LOP3_, SHF_, I2F_, FMUL_RZ, WRAP_FSET, WRAP_IMNMX, WRAP_PRED, N_OPS };
const char* names[] = {"FADD", "FFMA", "FADD2", "FFMA2", "FMUL2", "FSET.BF", "VIMNMX.U32", "FMNMX", "FSETP+FSEL", "FSETP+@P FADD", "MUFU.SIN", "MUFU.EX2",
                       "IMAD", "LOP3", "SHF", "I2FP", "FMUL.RZ", "wrap: FADD2+2xFSET+FADD2", "wrap: FADD2+FADD2+2xVIMNMX", "wrap: FADD2+2x(FSETP,@P FADD)"};
const int instr_per_step[] = {1, 1, 1, 1, 1, 1, 1, 1, 2, 2, 1, 1, 1, 1, 1, 1, 1, 4, 4, 5};

template <int OP, int C>
__global__ void k(float* out, long long* cyc, float seed) {
  float2 v[C];
  float2 a = make_float2(seed, seed * 1.5f), b = make_float2(0.999f, 1.001f);
#pragma unroll
  for (int c = 0; c < C; c++) v[c] = make_float2(seed + c, seed - c);
  __syncthreads();
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < ITERS; i++) {
#pragma unroll
    for (int u = 0; u < 4; u++) {
#pragma unroll
      for (int c = 0; c < C; c++) {
        float2& x = v[c];
        if (OP == FADD_) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(x.x) : "f"(a.x));
        if (OP == FFMA_) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x.x) : "f"(b.x), "f"(a.x));
        if (OP == FADD2_) x = __fadd2_rn(x, a);
        if (OP == FFMA2_) x = __ffma2_rn(x, b, a);
        if (OP == FMUL2_) x = __fmul2_rn(x, b);
        if (OP == FSETBF) asm volatile("set.ge.f32.f32 %0, %0, %1;" : "+f"(x.x) : "f"(a.x));
        if (OP == VIMNMX_) { unsigned r = __float_as_uint(x.x); asm volatile("min.u32 %0, %0, %1;" : "+r"(r) : "r"(__float_as_uint(a.x) + i)); x.x = __uint_as_float(r); }
        if (OP == FMNMX_) asm volatile("min.f32 %0, %0, %1;" : "+f"(x.x) : "f"(a.x));
        if (OP == FSETP_SEL) asm volatile("{.reg .pred p; setp.ge.f32 p, %0, %1; selp.f32 %0, %1, %2, p;}" : "+f"(x.x) : "f"(a.x), "f"(b.x));
        if (OP == FSETP_PADD) asm volatile("{.reg .pred p; setp.ge.f32 p, %0, %1; @p add.rn.f32 %0, %0, %2;}" : "+f"(x.x) : "f"(a.x), "f"(b.x));
        if (OP == MUFUSIN) asm volatile("sin.approx.ftz.f32 %0, %0;" : "+f"(x.x));
        if (OP == MUFUEX2) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x.x));
        if (OP == IMAD_) { int r = __float_as_int(x.x); asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(r) : "r"(__float_as_int(b.x)), "r"(i)); x.x = __int_as_float(r); }
        if (OP == LOP3_) { unsigned r = __float_as_uint(x.x); asm volatile("xor.b32 %0, %0, %1;" : "+r"(r) : "r"(__float_as_uint(a.x) + i)); x.x = __uint_as_float(r); }
        if (OP == SHF_) { unsigned r = __float_as_uint(x.x); asm volatile("shf.r.wrap.b32 %0, %0, %0, %1;" : "+r"(r) : "r"(i)); x.x = __uint_as_float(r); }
        if (OP == I2F_) { int r = __float_as_int(x.x); asm volatile("cvt.rn.f32.s32 %0, %1;" : "=f"(x.x) : "r"(r)); }
        if (OP == FMUL_RZ) asm volatile("mul.rz.f32 %0, %0, %1;" : "+f"(x.x) : "f"(b.x));
        if (OP == WRAP_FSET) {
          x = __fadd2_rn(x, a);
          float gx, gy;
          asm volatile("set.ge.f32.f32 %0, %1, 0f3F800000;" : "=f"(gx) : "f"(x.x));
          asm volatile("set.ge.f32.f32 %0, %1, 0f3F800000;" : "=f"(gy) : "f"(x.y));
          x = __fadd2_rn(x, make_float2(-gx, -gy));
        }
        if (OP == WRAP_IMNMX) {
          x = __fadd2_rn(x, a);
          float2 w = __fadd2_rn(x, make_float2(-1.f, -1.f));
          x = make_float2(__uint_as_float(min(__float_as_uint(x.x), __float_as_uint(w.x))), __uint_as_float(min(__float_as_uint(x.y), __float_as_uint(w.y))));
        }
        if (OP == WRAP_PRED) {
          x = __fadd2_rn(x, a);
          asm volatile("{.reg .pred p; setp.ge.f32 p, %0, 0f3F800000; @p add.rn.f32 %0, %0, 0fBF800000;}" : "+f"(x.x));
          asm volatile("{.reg .pred p; setp.ge.f32 p, %0, 0f3F800000; @p add.rn.f32 %0, %0, 0fBF800000;}" : "+f"(x.y));
        }
      }
    }
  }
  long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int c = 0; c < C; c++) s += v[c].x + v[c].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if ((threadIdx.x & 31) == 0) cyc[blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32] = t1 - t0;
}

template <int OP, int C>
void run(float* d_out, long long* d_cyc, int warps_per_smsp) {
  // one block per SM with 4 * W warps (one resident block per SM: 148 blocks)
  int threads = 32 * 4 * warps_per_smsp;
  k<OP, C><<<148, threads>>>(d_out, d_cyc, 0.37f);
  cudaDeviceSynchronize();
  k<OP, C><<<148, threads>>>(d_out, d_cyc, 0.37f);
  cudaDeviceSynchronize();
  static long long h[148 * 32];
  cudaMemcpy(h, d_cyc, sizeof(long long) * 148 * 4 * warps_per_smsp, cudaMemcpyDeviceToHost);
  long long mx = 0;
  for (int i = 0; i < 148 * 4 * warps_per_smsp; i++) mx = h[i] > mx ? h[i] : mx;
  // cycles per warp-STEP per SMSP = elapsed / (ITERS * 4 * C * W)
  double per_step = (double)mx / ((double)ITERS * 4.0 * C * warps_per_smsp);
  printf("%-30s chains=%d warps/SMSP=%d : %7.2f cycles per step per SMSP (%5.2f per instruction; one warp advances a step every %6.2f cycles)\n", names[OP], C,
         warps_per_smsp, per_step, per_step / instr_per_step[OP], (double)mx / ((double)ITERS * 4.0 * C) * C);
}

template <int OP>
void sweep(float* d_out, long long* d_cyc) {
  run<OP, 1>(d_out, d_cyc, 1);   // latency of the dependent chain
  run<OP, 8>(d_out, d_cyc, 1);   // one warp, 8 independent chains
  run<OP, 8>(d_out, d_cyc, 4);   // throughput
  run<OP, 8>(d_out, d_cyc, 8);
}

int main() {
  float* d_out; long long* d_cyc;
  cudaMalloc(&d_out, 148 * 1024 * 4);
  cudaMalloc(&d_cyc, 148 * 32 * 8);
  sweep<FADD_>(d_out, d_cyc); sweep<FFMA_>(d_out, d_cyc); sweep<FADD2_>(d_out, d_cyc); sweep<FFMA2_>(d_out, d_cyc); sweep<FMUL2_>(d_out, d_cyc);
  sweep<FSETBF>(d_out, d_cyc); sweep<VIMNMX_>(d_out, d_cyc); sweep<FMNMX_>(d_out, d_cyc); sweep<FSETP_SEL>(d_out, d_cyc); sweep<FSETP_PADD>(d_out, d_cyc);
  sweep<MUFUSIN>(d_out, d_cyc); sweep<MUFUEX2>(d_out, d_cyc); sweep<IMAD_>(d_out, d_cyc); sweep<LOP3_>(d_out, d_cyc); sweep<SHF_>(d_out, d_cyc);
  sweep<I2F_>(d_out, d_cyc); sweep<FMUL_RZ>(d_out, d_cyc);
  sweep<WRAP_FSET>(d_out, d_cyc); sweep<WRAP_IMNMX>(d_out, d_cyc); sweep<WRAP_PRED>(d_out, d_cyc);
  cudaError_t e = cudaGetLastError();
  printf("status: %s\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}
