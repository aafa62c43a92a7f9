// TMEM read micro-benchmark (sm_100a): is tensor memory usable as a per-lane constant-table store next to a
// shared-memory-bound kernel?
//
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/bin/ubench_tmem tools/ubench_tmem.cu && tools/bin/ubench_tmem
//
// One 512-thread CTA per SM.  Measured per SM:
//   (1) tcgen05.ld.32x32b.x{4,16,32} throughput with 4 / 8 / 16 warps issuing back to back
//   (2) LDS.128 throughput with the same warps (the shared-memory crossbar, 128 B/clk nominal)
//   (3) both together (8 warps each): do the two read paths share a port?
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

template <int X>
__device__ __forceinline__ uint32_t ldtm(uint32_t taddr);
template <>
__device__ __forceinline__ uint32_t ldtm<4>(uint32_t taddr) {
  uint32_t v0, v1, v2, v3;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v0), "=r"(v1), "=r"(v2), "=r"(v3) : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  return v0 ^ v1 ^ v2 ^ v3;
}
template <>
__device__ __forceinline__ uint32_t ldtm<16>(uint32_t taddr) {
  uint32_t v[16];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                 "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
               : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  uint32_t x = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) x ^= v[i];
  return x;
}
template <>
__device__ __forceinline__ uint32_t ldtm<32>(uint32_t taddr) {
  uint32_t v[32];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                 "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
                 "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
                 "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
               : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  uint32_t x = 0;
#pragma unroll
  for (int i = 0; i < 32; ++i) x ^= v[i];
  return x;
}

// mode bit 0: warps [0, n_tm) read TMEM; bit 1: warps [16 - n_lds, 16) read shared memory
template <int X>
__global__ void __launch_bounds__(512, 1) k_read(int n_tm, int n_lds, int iters, uint32_t* out, long long* cycles) {
  extern __shared__ __align__(16) float sm[];
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < 16384; i += 512) sm[i] = (float)i;
  if (warp == 0) {
    const uint32_t dst = (uint32_t)__cvta_generic_to_shared(&tmem_base_s);
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst), "r"(256) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_s;
  // fill the 256 columns of this warp's lane quadrant (warps 0-3) so that the loads return defined data
  if (warp < 4) {
    for (int c = 0; c < 256; c += 4) {
      const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c;
      const uint32_t a = lane * 256 + c;
      asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(taddr), "r"(a), "r"(a + 1), "r"(a + 2), "r"(a + 3) : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t acc = 0;
  const long long t0 = clock64();
  if (warp < n_tm) {
    const uint32_t tbase = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    for (int it = 0; it < iters; ++it) {
#pragma unroll 4
      for (int c = 0; c + X <= 256; c += X) acc ^= ldtm<X>(tbase + (uint32_t)c);
    }
  } else if (warp >= 16 - n_lds) {
    const unsigned sbase = (unsigned)__cvta_generic_to_shared(sm) + lane * 16;
    for (int it = 0; it < iters; ++it) {
#pragma unroll 8
      for (int c = 0; c < 64; ++c) {       // 64 LDS.128 = the same 32 KB per warp as 256 columns of 32 lanes
        float4 q;
        const unsigned ad = sbase + (((c + it) & 63) << 9);
        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(q.x), "=f"(q.y), "=f"(q.z), "=f"(q.w) : "r"(ad));
        acc ^= __float_as_uint(q.x) ^ __float_as_uint(q.y) ^ __float_as_uint(q.z) ^ __float_as_uint(q.w);
      }
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (tid == 0) cycles[blockIdx.x] = t1 - t0;
  out[blockIdx.x * 512 + tid] = acc;
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256) : "memory");
}

template <int X>
static int run(const char* name, int n_tm, int n_lds, uint32_t* out, long long* cyc) {
  const int iters = 200;
  CK(cudaFuncSetAttribute(k_read<X>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
  k_read<X><<<148, 512, 65536>>>(n_tm, n_lds, 10, out, cyc);
  CK(cudaDeviceSynchronize());
  k_read<X><<<148, 512, 65536>>>(n_tm, n_lds, iters, out, cyc);
  CK(cudaDeviceSynchronize());
  long long h[148];
  CK(cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost));
  double avg = 0;
  for (int i = 0; i < 148; ++i) avg += (double)h[i];
  avg /= 148.0;
  const double bytes_tm = (double)n_tm * 32.0 * 256.0 * 4.0 * iters;       // per SM
  const double bytes_ld = (double)n_lds * 64.0 * 512.0 * iters;
  printf("%-34s x%-2d tmem warps %2d lds warps %2d : %9.0f cycles  TMEM %7.1f B/clk/SM  LDS %7.1f B/clk/SM\n", name, X, n_tm, n_lds, avg,
         bytes_tm / avg, bytes_ld / avg);
  return 0;
}

int main() {
  uint32_t* out; long long* cyc;
  CK(cudaMalloc(&out, 148 * 512 * 4));
  CK(cudaMalloc(&cyc, 148 * 8));
  for (int w : {4, 8, 16}) { run<4>("tcgen05.ld only", w, 0, out, cyc); run<16>("tcgen05.ld only", w, 0, out, cyc); run<32>("tcgen05.ld only", w, 0, out, cyc); }
  for (int w : {4, 8, 16}) run<4>("LDS.128 only", 0, w, out, cyc);
  run<4>("tcgen05.ld + LDS.128 concurrently", 8, 8, out, cyc);
  run<16>("tcgen05.ld + LDS.128 concurrently", 8, 8, out, cyc);
  run<4>("tcgen05.ld alone (8 warps)", 8, 0, out, cyc);
  run<4>("LDS.128 alone (8 warps)", 0, 8, out, cyc);
  return 0;
}
