// micro-benchmark: FFMA vs FFMA2 (fma.rn.f32x2) issue/throughput on sm_100a, plus mixing with integer ops
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(float* out, int iters, float a, float b) {
  float x[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] = threadIdx.x * 0.001f + i;
  unsigned long long p[8];
  int z[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { asm("mov.b64 %0, {%1,%2};" : "=l"(p[i]) : "f"(x[2*i]), "f"(x[2*i+1])); z[i] = threadIdx.x + i; }
  unsigned long long ab, bb;
  asm("mov.b64 %0, {%1,%1};" : "=l"(ab) : "f"(a));
  asm("mov.b64 %0, {%1,%1};" : "=l"(bb) : "f"(b));
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {        // 16 scalar FFMA
#pragma unroll
      for (int i = 0; i < 16; ++i) x[i] = fmaf(x[i], a, b);
    } else if (MODE == 1) { // 8 FFMA2 (same flops)
#pragma unroll
      for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(ab), "l"(bb));
    } else if (MODE == 2) { // 16 FFMA + 8 integer LOP3/IADD
#pragma unroll
      for (int i = 0; i < 16; ++i) x[i] = fmaf(x[i], a, b);
#pragma unroll
      for (int i = 0; i < 8; ++i) z[i] = (z[i] ^ it) + i;
    } else if (MODE == 3) { // 8 FFMA2 + 8 integer
#pragma unroll
      for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(ab), "l"(bb));
#pragma unroll
      for (int i = 0; i < 8; ++i) z[i] = (z[i] ^ it) + i;
    } else if (MODE == 4) { // 16 FADD
#pragma unroll
      for (int i = 0; i < 16; ++i) x[i] = x[i] + a;
    } else if (MODE == 5) { // 8 FADD2
#pragma unroll
      for (int i = 0; i < 8; ++i) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(ab));
    }
  }
  float s = 0; int zs = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += x[i];
#pragma unroll
  for (int i = 0; i < 8; ++i) { float lo, hi; asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(p[i])); s += lo + hi; zs += z[i]; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + zs;
}
template <int MODE> void run(const char* name, float* out) {
  const int iters = 20000, blocks = 148 * 2, threads = 512;
  k<MODE><<<blocks, threads>>>(out, 100, 1.0001f, 0.5f);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<MODE><<<blocks, threads>>>(out, iters, 1.0001f, 0.5f);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double lane_ops = (double)iters * 16 * blocks * threads;   // fp32 lane-ops (fma or add)
  printf("%-28s %8.3f ms  %7.2f T lane-ops/s  (%.1f per clk per SM @1.965GHz)\n", name, ms, lane_ops / ms / 1e9, lane_ops / (ms * 1e-3) / 148 / 1.965e9);
}
int main() {
  float* out; cudaMalloc(&out, 148 * 2 * 512 * 4);
  run<0>("16 FFMA", out); run<1>("8 FFMA2", out); run<2>("16 FFMA + 8 int", out); run<3>("8 FFMA2 + 8 int", out);
  run<4>("16 FADD", out); run<5>("8 FADD2", out);
  cudaDeviceSynchronize(); printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
