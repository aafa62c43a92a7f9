// One 32-point DFT pass of the Griffin-Lim frame transform on the tensor core vs on the FP32 pipe (sm_100a).
//
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -I your-voice-tts_b200/csrc \
//        -o tools/bin/ubench_dft_tc tools/ubench_dft_tc.cu && tools/bin/ubench_dft_tc
//
// The frame kernels transform 1024 packed complex points as 32 x 32: every lane of a warp runs a 32-point complex DFT
// on 32 values it holds in registers (fft32p.cuh), twice per transform, four times per frame and iteration.  This
// program measures the alternative VERDICT r1 asked for: the same pass as a GEMM on tcgen05,
//
//     D[128 x 64] = A[128 x 64] * B[64 x 64]^T     (rows = the lanes of 4 warps = 4 frames, 64 = re/im of 32 points)
//
// with fp32-class accuracy from a 2-term bf16 split (A = Ah + Al, B = Bh + Bl; Ah Bh + Ah Bl + Al Bh = 3 x 4 MMAs of
// shape 128 x 64 x 16).  A goes registers -> (cvt, sub, cvt) -> tcgen05.st -> TENSOR MEMORY (no shared-memory traffic
// for the data), B (the DFT matrix, pre-split, 2 x 8 KB) stays resident in shared memory, D comes back with
// tcgen05.ld.32x32b (TMEM lane = thread: the lane gets its own 64 outputs, exactly what fft32p returns).
//
// One 512-thread CTA per SM (the geometry of gl_stream_kernel): 4 groups of 4 warps, each group owns 128 TMEM columns
// (Ah 32 | Al 32 | D 64).  Modes:
//   fp32     : every pass is fft32p                                   (the shipped arithmetic)
//   tc       : every pass on the tensor core
//   mixed    : passes alternate fp32 / tc                             (2 of 4 passes moved, both pipes busy)
//   prep     : operand preparation + tcgen05.st only, no MMA          (what the CUDA cores pay per tensor pass)
//   mma      : MMAs + commit + wait + tcgen05.ld only, no conversion  (what the tensor pipe and the round trip cost)
// Reported: cycles per frame-pass per SM (16 warps resident), accuracy of ONE pass against a float64 DFT.
#include <cstdio>
#include <cstdint>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include "fft32p.cuh"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); return 1; } } while (0)

using namespace ttsa;

constexpr int kWarps = 16, kThreads = 512;
constexpr uint32_t kLboB = (64 / 8) * 128;      // canonical K-major, no swizzle: byte stride between 8-element K blocks
constexpr uint32_t kBBytes = 64 * 64 * 2;
// instruction descriptor: D = f32, A = B = bf16, K-major, N = 64, M = 128
constexpr uint32_t kIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

enum { MODE_FP32 = 0, MODE_TC = 1, MODE_MIXED = 2, MODE_PREP = 3, MODE_MMA = 4 };

__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
               ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ uint32_t pack_bf16(float lo_half, float hi_half) {   // result[15:0] = bf16(lo_half)
  uint32_t d;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi_half), "f"(lo_half));
  return d;
}
__device__ __forceinline__ void st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
               ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
                 "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
}
__device__ __forceinline__ void ld32(uint32_t taddr, float (&v)[32]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
               : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]), "=f"(v[8]), "=f"(v[9]),
                 "=f"(v[10]), "=f"(v[11]), "=f"(v[12]), "=f"(v[13]), "=f"(v[14]), "=f"(v[15]), "=f"(v[16]), "=f"(v[17]), "=f"(v[18]), "=f"(v[19]),
                 "=f"(v[20]), "=f"(v[21]), "=f"(v[22]), "=f"(v[23]), "=f"(v[24]), "=f"(v[25]), "=f"(v[26]), "=f"(v[27]), "=f"(v[28]), "=f"(v[29]),
                 "=f"(v[30]), "=f"(v[31])
               : "r"(taddr) : "memory");
}

// data: [rows][32] complex (float2), one row per thread; out the same.  bimg: B hi | B lo in the canonical layout.
__global__ void __launch_bounds__(kThreads, 1)
k_pass(int mode, int iters, const float2* __restrict__ in, float2* __restrict__ out, const uint4* __restrict__ bimg, long long* cycles) {
  __shared__ __align__(1024) unsigned char smB[2 * kBBytes];
  __shared__ __align__(8) uint64_t mbar[4];
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, grp = warp >> 2;

  for (int i = tid; i < (int)(2 * kBBytes / 16); i += kThreads) reinterpret_cast<uint4*>(smB)[i] = bimg[i];
  if (warp == 0) {
    const uint32_t dst = (uint32_t)__cvta_generic_to_shared(&tmem_base_s);
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid < 4) {
    const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&mbar[tid]);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)grp * 128u;   // this warp's lanes, its group's columns
  const uint32_t tmem_grp = tmem_base_s + (uint32_t)grp * 128u;
  const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&mbar[grp]);
  const uint32_t smB_addr = (uint32_t)__cvta_generic_to_shared(smB);

  float2 R[16], I[16];
  const size_t row = (size_t)blockIdx.x * kThreads + tid;
#pragma unroll
  for (int m = 0; m < 16; ++m) {
    const float2 a = in[row * 32 + 2 * m], b = in[row * 32 + 2 * m + 1];
    R[m] = make_float2(a.x, b.x);
    I[m] = make_float2(a.y, b.y);
  }
  uint32_t parity = 0;
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
    const bool tc = mode == MODE_TC || mode == MODE_PREP || mode == MODE_MMA || (mode == MODE_MIXED && (it & 1));
    if (!tc) {
      fft32p(R, I);
      constexpr float s = 0.17677669529663687f;                    // unitary scaling keeps repeated passes bounded
#pragma unroll
      for (int m = 0; m < 16; ++m) { R[m] = __fmul2_rn(R[m], splat(s)); I[m] = __fmul2_rn(I[m], splat(s)); }
    } else {
      if (mode != MODE_MMA) {
        // ---- operand preparation: x = hi + lo, both bf16 pairs; K order = re[0..31] | im[0..31]
        // (two halves of 16 columns each, so that at most 32 packed values are live next to the 64 data registers)
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          uint32_t hi[16], lo[16];
#pragma unroll
          for (int m = 0; m < 16; ++m) {
            const float2 x = half ? I[m] : R[m];
            hi[m] = pack_bf16(x.x, x.y);
            const float2 h = make_float2(__uint_as_float(hi[m] << 16), __uint_as_float(hi[m] & 0xffff0000u));
            const float2 l = __fadd2_rn(x, neg2(h));
            lo[m] = pack_bf16(l.x, l.y);
          }
          st16(tmem + 16 * half, hi);
          st16(tmem + 32 + 16 * half, lo);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      }
      if (mode != MODE_PREP) {
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        asm volatile("bar.sync %0, 128;" ::"r"(grp + 1) : "memory");
        if ((warp & 3) == 0 && lane == 0) {
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
          for (int term = 0; term < 3; ++term) {                   // Ah Bh, Ah Bl, Al Bh
            const uint32_t a_col = term == 2 ? 32u : 0u;
            const uint32_t b_off = term == 1 ? kBBytes : 0u;
#pragma unroll
            for (int j = 0; j < 4; ++j)
              mma_ts(tmem_grp + 64, tmem_grp + a_col + 8 * j, smem_desc(smB_addr + b_off + 2 * j * kLboB, kLboB, 128), kIdesc, (term | j) != 0);
          }
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
        }
        uint32_t done = 0;
        while (!done) {
          asm volatile("{\n\t.reg .pred q;\n\tmbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2;\n\tselp.u32 %0, 1, 0, q;\n\t}\n"
                       : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        }
        parity ^= 1;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        float vr[32], vi[32];
        ld32(tmem + 64, vr);
        ld32(tmem + 96, vi);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int m = 0; m < 16; ++m) { R[m] = make_float2(vr[2 * m], vr[2 * m + 1]); I[m] = make_float2(vi[2 * m], vi[2 * m + 1]); }
      } else {
        R[0].x += 1e-30f;      // keep the loop body dependent
      }
    }
  }
  const long long t1 = clock64();
  __syncthreads();
  if (tid == 0) cycles[blockIdx.x] = t1 - t0;
#pragma unroll
  for (int m = 0; m < 16; ++m) {
    out[row * 32 + 2 * m] = make_float2(R[m].x, I[m].x);
    out[row * 32 + 2 * m + 1] = make_float2(R[m].y, I[m].y);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base_s), "r"(512) : "memory");
}

static uint16_t f2bf(float x) {   // round to nearest even
  uint32_t u; memcpy(&u, &x, 4);
  u += 0x7fffu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}
static float bf2f(uint16_t h) { uint32_t u = (uint32_t)h << 16; float f; memcpy(&f, &u, 4); return f; }

int main() {
  const int grid = 148, rows = grid * kThreads;
  std::vector<float2> h_in((size_t)rows * 32), h_out((size_t)rows * 32);
  uint64_t s = 0x9E3779B97F4A7C15ull;
  auto rnd = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return (float)((double)(s >> 11) / 9007199254740992.0 * 2.0 - 1.0); };
  for (auto& v : h_in) v = make_float2(rnd(), rnd());
  // B[n][k]: out re j = sum_i re_i cos + im_i sin;  out im j = sum_i im_i cos - re_i sin;  unitary scale 1/sqrt(32)
  std::vector<uint16_t> h_b(2 * 64 * 64);
  const double sc = 1.0 / std::sqrt(32.0);
  for (int n = 0; n < 64; ++n)
    for (int k = 0; k < 64; ++k) {
      const int j = n & 31, i = k & 31;
      const double th = 2.0 * M_PI * (double)((i * j) & 31) / 32.0;
      double v;
      if (n < 32) v = (k < 32) ? std::cos(th) : std::sin(th);
      else v = (k < 32) ? -std::sin(th) : std::cos(th);
      v *= sc;
      const uint16_t h = f2bf((float)v);
      const uint16_t l = f2bf((float)(v - (double)bf2f(h)));
      const size_t off = (size_t)(k >> 3) * (kLboB / 2) + (size_t)(n >> 3) * 64 + (size_t)(n & 7) * 8 + (size_t)(k & 7);
      h_b[off] = h;
      h_b[64 * 64 + off] = l;
    }
  float2 *d_in, *d_out; uint4* d_b; long long* d_cyc;
  CK(cudaMalloc(&d_in, h_in.size() * 8)); CK(cudaMalloc(&d_out, h_in.size() * 8));
  CK(cudaMalloc(&d_b, h_b.size() * 2)); CK(cudaMalloc(&d_cyc, grid * 8));
  CK(cudaMemcpy(d_in, h_in.data(), h_in.size() * 8, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_b, h_b.data(), h_b.size() * 2, cudaMemcpyHostToDevice));

  // ---- accuracy of one pass
  for (int mode : {MODE_FP32, MODE_TC}) {
    k_pass<<<grid, kThreads>>>(mode, 1, d_in, d_out, d_b, d_cyc);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(h_out.data(), d_out, h_out.size() * 8, cudaMemcpyDeviceToHost));
    double num = 0, den = 0, worst = 0;
    for (int r = 0; r < 4096; ++r) {
      const int rr = (int)(((long long)r * 2654435761ll) % rows);
      for (int k = 0; k < 32; ++k) {
        double xr = 0, xi = 0;
        for (int j = 0; j < 32; ++j) {
          const double th = -2.0 * M_PI * (double)((j * k) & 31) / 32.0;
          const double a = h_in[(size_t)rr * 32 + j].x, b = h_in[(size_t)rr * 32 + j].y;
          xr += a * std::cos(th) - b * std::sin(th);
          xi += a * std::sin(th) + b * std::cos(th);
        }
        xr *= sc; xi *= sc;
        const double er = h_out[(size_t)rr * 32 + k].x - xr, ei = h_out[(size_t)rr * 32 + k].y - xi;
        num += er * er + ei * ei; den += xr * xr + xi * xi;
        worst = std::max(worst, std::sqrt(er * er + ei * ei));
      }
    }
    printf("accuracy %-5s one pass vs float64 DFT: SNR %.1f dB, worst |err| %.2e (values ~ unit variance)\n",
           mode == MODE_FP32 ? "fp32" : "tc", 10.0 * std::log10(den / num), worst);
  }
  // ---- timing
  const char* names[5] = {"fp32 (fft32p + scale)", "tc (prep + 12 MMA + ld)", "mixed (alternate)", "prep only (cvt/sub/cvt + st)", "mma only (12 MMA + ld)"};
  const int iters = 2000;
  for (int mode = 0; mode < 5; ++mode) {
    k_pass<<<grid, kThreads>>>(mode, 50, d_in, d_out, d_b, d_cyc);
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k_pass<<<grid, kThreads>>>(mode, iters, d_in, d_out, d_b, d_cyc);
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    std::vector<long long> h(grid);
    CK(cudaMemcpy(h.data(), d_cyc, grid * 8, cudaMemcpyDeviceToHost));
    double avg = 0; for (auto c : h) avg += (double)c; avg /= grid;
    printf("%-30s %8.1f cycles per pass of 16 warps = %6.1f cycles per frame-pass per SM   (%.3f ms for %d passes)\n",
           names[mode], avg / iters, avg / iters / 16.0, ms, iters);
  }
  return 0;
}
