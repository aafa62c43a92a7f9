// mel <-> linear as a tensor-core GEMM (tcgen05.mma, accumulator in TMEM) with fp32-equivalent precision.
//
//   _mel_to_linear   lin[row, k] = max(1e-10, sum_m pinv[k, m] amp[row, m])      utils/audio.py:64-66
//   _linear_to_mel   mel[row, m] = sum_k basis[m, k] amp[row, k]                 utils/audio.py:60-62
//
// Both are D[rows x N] = A[rows x K] * B[N x K]^T with a plan-constant B (pseudo-inverse: N = 1025, K = 80; mel basis:
// N = 80, K = 1025).  A single-pass TF32/BF16 product fails the 60 dB Griffin-Lim bar on the sign-cancelling
// pseudo-inverse (SURVEY: 40 dB), so each fp32 operand is split into three bf16 terms (hi + mid + lo = 24 mantissa
// bits) and the six products hi*hi, hi*mid, mid*hi, hi*lo, mid*mid, lo*hi are accumulated in fp32 in tensor memory:
// 6 x (K/16) UMMA instructions of shape 128 x N x 16 per tile.
//
// CTA = 256 threads, one 128-row tile of A, one N-tile of B, K walked in chunks of 80.  Per chunk: A is converted
// (prologue: denormalise + dB->amplitude when the input is a normalised spectrogram), split and written to shared
// memory in the canonical K-major no-swizzle UMMA layout (8-row x 16-byte core matrices); B was pre-split and
// pre-arranged in that layout at plan creation and is copied verbatim.  One elected thread issues the MMAs and commits
// them to an mbarrier.  Epilogue: tcgen05.ld (one TMEM lane = one output row per thread), clamp / power / dB, staged
// through shared memory for coalesced row-major stores.
#pragma once
#include <cuda_bf16.h>
#include "aux_kernels.cuh"

namespace ttsa {

constexpr int kTcRows = 128;        // UMMA M
constexpr int kTcChunk = 80;        // K per chunk (5 UMMA K-steps of 16 bf16)
constexpr int kTcKSteps = kTcChunk / 16;
constexpr int kTcThreads = 256;     // 8 warps: warps w and w + 4 share TMEM lanes 32 (w % 4) .. + 31 and split the columns

struct TcGemmParams {
  MelParams mp;
  const float* a;            // [rows, lda] fp32
  int lda;                   // elements per A row (80 or 1025)
  int k_total;               // valid K (A columns); chunks beyond are zero padded
  int n_chunks;
  const __nv_bfloat16* b;    // canonical pre-split B: [n_tile][chunk][part 3][kb 10][rg N/8][8][8]
  float* out;                // [rows, ldo]
  int ldo;                   // 1025 or num_mels
  int n_valid;               // valid output columns (1025 or num_mels)
  int in_kind, out_kind;
  int mode;                  // 0 = mel_to_linear epilogue (max 1e-10, optional power), 1 = linear_to_mel (optional norm dB)
};

__device__ __forceinline__ uint64_t umma_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  // K-major, SWIZZLE_NONE: start address, leading (K) byte offset, stride (M/N) byte offset in 16-byte units; version 1
  return (uint64_t)((smem_addr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46);
}

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

template <int N_TILE>   // UMMA N (multiple of 16, <= 256)
__global__ void __launch_bounds__(kTcThreads, 1) gemm_bf16x3_tc_kernel(const TcGemmParams p) {
  extern __shared__ __align__(1024) unsigned char tc_smem[];
  constexpr int kTmemCols = N_TILE <= 32 ? 32 : N_TILE <= 64 ? 64 : N_TILE <= 128 ? 128 : 256;
  constexpr uint32_t kABytes = kTcRows * kTcChunk * 2;        // one bf16 part of the A chunk
  constexpr uint32_t kBBytes = N_TILE * kTcChunk * 2;         // one bf16 part of the B chunk
  constexpr uint32_t kLboA = (kTcRows / 8) * 128;             // byte stride between 8-element K blocks
  constexpr uint32_t kLboB = (N_TILE / 8) * 128;
  constexpr int kStageLd = N_TILE + 1;
  // instruction descriptor: D = f32, A = B = bf16, both K-major, N >> 3 at bit 17, M >> 4 at bit 24
  constexpr uint32_t kIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N_TILE >> 3) << 17) | ((uint32_t)(kTcRows >> 4) << 24);

  unsigned char* const smA = tc_smem;                          // 3 parts
  unsigned char* const smB = tc_smem + 3 * kABytes;            // 3 parts
  float* const stage = reinterpret_cast<float*>(tc_smem);      // epilogue staging (operands are dead by then)
  __shared__ __align__(8) uint64_t mbar;
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long row0 = (long long)blockIdx.x * kTcRows;
  const int n_tile = blockIdx.y;

  if (warp == 0) {
    const uint32_t dst = (uint32_t)__cvta_generic_to_shared(&tmem_base_s);
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst), "r"(kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&mbar);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_s;
  const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&mbar);
  const uint32_t smA_addr = (uint32_t)__cvta_generic_to_shared(smA);
  const uint32_t smB_addr = (uint32_t)__cvta_generic_to_shared(smB);

  for (int ch = 0; ch < p.n_chunks; ++ch) {
    if (ch > 0) {
      // the previous chunk's MMAs must have consumed shared memory before it is overwritten
      const uint32_t parity = (uint32_t)((ch - 1) & 1);
      uint32_t done = 0;
      while (!done) {
        asm volatile(
            "{\n\t.reg .pred q;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, q;\n\t}\n"
            : "=r"(done) : "r"(bar), "r"(parity) : "memory");
      }
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    // ---- A chunk: fp32 -> amplitude -> three bf16 terms, canonical layout
    const int k0 = ch * kTcChunk;
    for (int e = tid; e < kTcRows * kTcChunk; e += kTcThreads) {
      const int r = e / kTcChunk, kk = e - r * kTcChunk;
      const long long row = row0 + r;
      const int k = k0 + kk;
      float x = 0.0f;
      if (row < p.mp.rows && k < p.k_total) {
        x = mel_in_value(__ldg(p.a + row * p.lda + k), p.in_kind, p.mp);
        if (p.mode == 1 && p.in_kind == 1) x = fabsf(x);       // out_linear_to_mel takes np.abs (utils/audio.py:177)
      }
      const __nv_bfloat16 h = __float2bfloat16_rn(x);
      const float r1 = x - __bfloat162float(h);
      const __nv_bfloat16 m = __float2bfloat16_rn(r1);
      const __nv_bfloat16 l = __float2bfloat16_rn(r1 - __bfloat162float(m));
      const uint32_t off = (uint32_t)(kk >> 3) * kLboA + (uint32_t)(r >> 3) * 128u + (uint32_t)(r & 7) * 16u + (uint32_t)(kk & 7) * 2u;
      *reinterpret_cast<__nv_bfloat16*>(smA + off) = h;
      *reinterpret_cast<__nv_bfloat16*>(smA + kABytes + off) = m;
      *reinterpret_cast<__nv_bfloat16*>(smA + 2 * kABytes + off) = l;
    }
    // ---- B chunk: already split and laid out; verbatim 16-byte copies
    {
      const uint4* src = reinterpret_cast<const uint4*>(p.b) + ((size_t)n_tile * p.n_chunks + ch) * (3 * kBBytes / 16);
      uint4* dst = reinterpret_cast<uint4*>(smB);
      for (int i = tid; i < (int)(3 * kBBytes / 16); i += kTcThreads) dst[i] = __ldg(src + i);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy writes -> visible to the tensor core
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (tid == 0) {
      // (A part, B part): hi*hi, hi*mid, mid*hi, hi*lo, mid*mid, lo*hi
      const int pa[6] = {0, 0, 1, 0, 1, 2};
      const int pb[6] = {0, 1, 0, 2, 1, 0};
#pragma unroll
      for (int c = 0; c < 6; ++c) {
#pragma unroll
        for (int j = 0; j < kTcKSteps; ++j) {
          const uint64_t da = umma_smem_desc(smA_addr + pa[c] * kABytes + 2 * j * kLboA, kLboA, 128);
          const uint64_t db = umma_smem_desc(smB_addr + pb[c] * kBBytes + 2 * j * kLboB, kLboB, 128);
          umma_bf16(tmem_base, da, db, kIdesc, (ch | c | j) != 0 ? 1u : 0u);
        }
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
    }
  }
  // ---- wait for the last chunk's MMAs
  {
    const uint32_t parity = (uint32_t)((p.n_chunks - 1) & 1);
    uint32_t done = 0;
    while (!done) {
      asm volatile(
          "{\n\t.reg .pred q;\n\t"
          "mbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2;\n\t"
          "selp.u32 %0, 1, 0, q;\n\t}\n"
          : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }
  __syncthreads();                                    // every thread is past its wait: operands may be overwritten

  // ---- epilogue: TMEM lane = output row (warp w may touch lanes 32 (w % 4) .. + 31), 16 columns per load;
  //      warps 0-3 take the first half of the column chunks, warps 4-7 the second
  {
    const int lane_base = (warp & 3) * 32;
    const int r = lane_base + lane;
    constexpr int kChunks = N_TILE / 16;
    const int c_begin = (warp < 4) ? 0 : (kChunks + 1) / 2;
    const int c_end = (warp < 4) ? (kChunks + 1) / 2 : kChunks;
    const bool pow15 = p.mp.power == 1.5f;
#pragma unroll 1
    for (int cc = c_begin; cc < c_end; ++cc) {
      const int c0 = cc * 16;
      uint32_t v[16];
      const uint32_t taddr = tmem_base + ((uint32_t)lane_base << 16) + (uint32_t)c0;
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
          : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
            "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
          : "r"(taddr) : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        float x = __uint_as_float(v[i]);
        if (p.mode == 0) {
          x = fmaxf(1e-10f, x);
          if (p.out_kind == 1) {
          if (pow15) {                         // x ** 1.5 (the shipped `power`): x * sqrt(x), sqrt.approx is ~1 ulp
            float sq;
            asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(sq) : "f"(x));
            x *= sq;
          } else {
            x = exp2f(p.mp.power * log2f(x));
          }
        }
        } else if (p.out_kind == 2) {
          x = fmaf(p.mp.n_a, log2f(fmaxf(p.mp.min_amp, x)), p.mp.n_b);
          x = fminf(fmaxf(x, p.mp.n_lo), p.mp.n_hi);
        }
        stage[r * kStageLd + c0 + i] = x;
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  // ---- coalesced row-major stores
  {
    const int col0 = n_tile * N_TILE;
    for (int r = warp; r < kTcRows; r += kTcThreads / 32) {
      const long long row = row0 + r;
      if (row >= p.mp.rows) break;
      float* dst = p.out + row * p.ldo + col0;
      for (int c = lane; c < N_TILE; c += 32)
        if (col0 + c < p.n_valid) dst[c] = stage[r * kStageLd + c];
    }
  }
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
  }
}

}  // namespace ttsa

// ---------------------------------------------------------------------------------------------------------
// Pipelined mel -> linear (num_mels <= 80): one CTA walks the N tiles of a 128-frame M tile.
//   A (mel amplitudes, 3 bf16 terms) is converted once and stays in shared memory;
//   B tiles (pseudo-inverse rows) stream in with cp.async into two alternating buffers;
//   the accumulator alternates between two TMEM regions, so the MMAs of N tile nt run while tile nt-1 drains:
//   tcgen05.ld -> clamp/power -> shared-memory staging -> coalesced row stores.
// ---------------------------------------------------------------------------------------------------------
namespace ttsa {

constexpr int kM2lN = 96;                         // UMMA N of the pipelined kernel; 11 tiles cover 1056 >= 1025 bins
constexpr int kM2lTiles = 11;
constexpr int kM2lThreads = 512;
constexpr int kM2lStageLd = kM2lN + 1;
constexpr uint32_t kM2lABytes = kTcRows * kTcChunk * 2;          // 20 480 per part
constexpr uint32_t kM2lBBytes = kM2lN * kTcChunk * 2;            // 15 360 per part
constexpr size_t kM2lSmem = 3 * kM2lABytes + 2 * 3 * kM2lBBytes + (size_t)kTcRows * kM2lStageLd * 4;

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred q;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, q;\n\t}\n"
        : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  }
}

__global__ void __launch_bounds__(kM2lThreads, 1) mel_to_linear_tc_kernel(const TcGemmParams p, const int n_mtiles) {
  extern __shared__ __align__(1024) unsigned char tc_smem[];
  constexpr uint32_t kLboA = (kTcRows / 8) * 128, kLboB = (kM2lN / 8) * 128;
  constexpr uint32_t kIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kM2lN >> 3) << 17) | ((uint32_t)(kTcRows >> 4) << 24);
  unsigned char* const smA = tc_smem;
  unsigned char* const smB = tc_smem + 3 * kM2lABytes;                       // two buffers of 3 parts
  float* const stage = reinterpret_cast<float*>(tc_smem + 3 * kM2lABytes + 2 * 3 * kM2lBBytes);
  __shared__ __align__(8) uint64_t mbar[2];
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (warp == 0) {
    const uint32_t dst = (uint32_t)__cvta_generic_to_shared(&tmem_base_s);
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst), "r"(256) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((uint32_t)__cvta_generic_to_shared(&mbar[0])) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((uint32_t)__cvta_generic_to_shared(&mbar[1])) : "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_s;
  const uint32_t bar0 = (uint32_t)__cvta_generic_to_shared(&mbar[0]);
  const uint32_t smA_addr = (uint32_t)__cvta_generic_to_shared(smA);
  const uint32_t smB_addr = (uint32_t)__cvta_generic_to_shared(smB);
  const bool pow15 = p.mp.power == 1.5f;
  uint32_t uses[2] = {0, 0};                       // completed uses of each accumulator / mbarrier (phase parity)

  auto load_b = [&](int nt) {                      // pseudo-inverse rows of N tile nt -> buffer nt & 1 (async)
    const uint4* src = reinterpret_cast<const uint4*>(p.b) + (size_t)nt * (3 * kM2lBBytes / 16);
    unsigned char* dst = smB + (nt & 1) * 3 * kM2lBBytes;
    for (int i = tid; i < (int)(3 * kM2lBBytes / 16); i += kM2lThreads) {
      const unsigned d = (unsigned)__cvta_generic_to_shared(dst + 16 * i);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src + i) : "memory");
    }
  };
  auto epilogue_wait = [&](int nt) {               // the MMAs of tile nt are complete: accumulator nt & 1 is readable and
    const int b = nt & 1;                          // the B buffer nt & 1 may be refilled
    mbar_wait(bar0 + 8 * b, uses[b] & 1);
    uses[b] += 1;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  };
  auto epilogue = [&](int nt, long long row0) {    // drain accumulator nt & 1 (after epilogue_wait(nt))
    const int b = nt & 1;
    const int lane_base = (warp & 3) * 32;
    const int r = lane_base + lane;
#pragma unroll 1
    for (int cc = (warp >> 2); cc < kM2lN / 8; cc += kM2lThreads / 128) {     // 12 chunks of 8 columns over 4 warp groups
      uint32_t v[8];
      const uint32_t taddr = tmem_base + ((uint32_t)lane_base << 16) + (uint32_t)(b * kM2lN + cc * 8);
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                   : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                   : "r"(taddr) : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float x = fmaxf(1e-10f, __uint_as_float(v[i]));
        if (p.out_kind == 1) {
          if (pow15) {                         // x ** 1.5 (the shipped `power`): x * sqrt(x), sqrt.approx is ~1 ulp
            float sq;
            asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(sq) : "f"(x));
            x *= sq;
          } else {
            x = exp2f(p.mp.power * log2f(x));
          }
        }
        stage[r * kM2lStageLd + cc * 8 + i] = x;
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    const int col0 = nt * kM2lN;
    // warp w stores rows w, w + 16, ...; its lanes cover the 96 columns in three 128-byte pieces
    static_assert(kM2lN == 96 && kM2lThreads == 512, "store mapping below assumes 96 columns and 16 warps");
    for (int rr = warp; rr < kTcRows; rr += kM2lThreads / 32) {
      const long long row = row0 + rr;
      if (row >= p.mp.rows) break;
      float* orow = p.out + row * p.ldo + col0;
      const float* srow = stage + rr * kM2lStageLd;
#pragma unroll
      for (int c = lane; c < kM2lN; c += 32)
        if (col0 + c < p.n_valid) orow[c] = srow[c];
    }
  };

  for (int mt = blockIdx.x; mt < n_mtiles; mt += gridDim.x) {
    const long long row0 = (long long)mt * kTcRows;
    __syncthreads();                                // previous M tile fully drained (staging, A, B reusable)
    load_b(0);
    // ---- A: fp32 mel -> amplitude -> three bf16 terms; thread (row, quarter) converts 3 / 3 / 2 / 2 K blocks of 8
    {
      const int r = tid >> 2, quarter = tid & 3;
      const long long row = row0 + r;
      const int kb0 = quarter < 2 ? 3 * quarter : 6 + 2 * (quarter - 2), kb1 = kb0 + (quarter < 2 ? 3 : 2);
#pragma unroll 1
      for (int kb = kb0; kb < kb1; ++kb) {
        __align__(16) __nv_bfloat16 h8[8], m8[8], l8[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int k = kb * 8 + i;
          float x = 0.0f;
          if (row < p.mp.rows && k < p.k_total) x = mel_in_value(__ldg(p.a + row * p.lda + k), p.in_kind, p.mp);
          h8[i] = __float2bfloat16_rn(x);
          const float r1 = x - __bfloat162float(h8[i]);
          m8[i] = __float2bfloat16_rn(r1);
          l8[i] = __float2bfloat16_rn(r1 - __bfloat162float(m8[i]));
        }
        const uint32_t off = (uint32_t)kb * kLboA + (uint32_t)(r >> 3) * 128u + (uint32_t)(r & 7) * 16u;
        *reinterpret_cast<uint4*>(smA + off) = *reinterpret_cast<const uint4*>(h8);
        *reinterpret_cast<uint4*>(smA + kM2lABytes + off) = *reinterpret_cast<const uint4*>(m8);
        *reinterpret_cast<uint4*>(smA + 2 * kM2lABytes + off) = *reinterpret_cast<const uint4*>(l8);
      }
    }
    for (int nt = 0; nt < kM2lTiles; ++nt) {
      asm volatile("cp.async.wait_all;" ::: "memory");                 // B(nt) has landed (this thread's part)
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy writes -> visible to the tensor core
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncthreads();
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (tid == 0) {
        const int pa[6] = {0, 0, 1, 0, 1, 2};
        const int pb[6] = {0, 1, 0, 2, 1, 0};
        const uint32_t bbase = smB_addr + (nt & 1) * 3 * kM2lBBytes;
        const uint32_t dcol = tmem_base + (uint32_t)((nt & 1) * kM2lN);
#pragma unroll
        for (int c = 0; c < 6; ++c) {
#pragma unroll
          for (int j = 0; j < kTcKSteps; ++j) {
            const uint64_t da = umma_smem_desc(smA_addr + pa[c] * kM2lABytes + 2 * j * kLboA, kLboA, 128);
            const uint64_t db = umma_smem_desc(bbase + pb[c] * kM2lBBytes + 2 * j * kLboB, kLboB, 128);
            umma_bf16(dcol, da, db, kIdesc, (c | j) != 0 ? 1u : 0u);
          }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar0 + 8 * (nt & 1)) : "memory");
      }
      // the other B buffer was last read by the MMAs of tile nt-1: once they are complete, the copy of tile nt+1 is issued
      // into it BEFORE the epilogue work of tile nt-1, so that it streams in behind that work and the MMAs of tile nt
      if (nt >= 1) epilogue_wait(nt - 1);
      if (nt + 1 < kM2lTiles) load_b(nt + 1);
      if (nt >= 1) epilogue(nt - 1, row0);
    }
    __syncthreads();                                // the row stores of tile kM2lTiles-2 have read the staging buffer
    epilogue_wait(kM2lTiles - 1);
    epilogue(kM2lTiles - 1, row0);
  }
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256) : "memory");
  }
}

}  // namespace ttsa

// ---------------------------------------------------------------------------------------------------------
// mel -> linear, transposed and warp-specialised (num_mels <= 80): D^T[bins x frames] = pinv[bins x K] * amp[frames x K]^T.
//   With the BINS as the UMMA M dimension a TMEM lane is a bin and a column a frame, so after tcgen05.ld (one lane per
//   thread) the 32 lanes of a warp hold 32 consecutive bins of one frame in each register: every register goes to HBM
//   with one coalesced 128-byte store, straight from registers -- no shared-memory staging and no CTA barrier in the
//   epilogue (the kernel above spends 3.4 us per 128 x 96 tile there, where the 30 MMAs need 0.7 us).
//   16 epilogue warps (TMEM lane group = warp % 4, 32 frames each = warp / 4) + 1 producer warp whose lane 0 streams the
//   pre-split pseudo-inverse tiles (one 60 KB bulk copy per 128-bin tile, two buffers) and issues the MMAs into two
//   alternating TMEM accumulators.  mbarriers: tile landed (tx bytes), accumulator full (tcgen05.commit), accumulator
//   drained (one arrival per epilogue warp, right after its tcgen05.ld).  The frame operand (128 frames x 80 mels, three
//   bf16 terms) is converted once per frame tile by the epilogue warps.
// ---------------------------------------------------------------------------------------------------------
namespace ttsa {

constexpr int kT2Bins = 128;                       // UMMA M: bins per tile
constexpr int kT2Tiles = 9;                        // 1152 >= 1025 bins
constexpr int kT2Frames = 128;                     // UMMA N: frames per tile
constexpr int kT2EpiWarps = 16;
constexpr int kT2Threads = (kT2EpiWarps + 1) * 32; // + the producer warp
constexpr uint32_t kT2PartBytes = 128 * kTcChunk * 2;            // one bf16 term of a 128-row operand: 20 480
constexpr uint32_t kT2TileBytes = 3 * kT2PartBytes;              // 61 440
constexpr size_t kT2Smem = (size_t)kT2TileBytes * 3 + 1024;      // frames + two pinv buffers (+ alignment slack)
#ifndef TTSA_T2_PIECES
#define TTSA_T2_PIECES 12
#endif
constexpr int kT2Pieces = TTSA_T2_PIECES;                        // bulk copies per pseudo-inverse tile (5 KB each)
static_assert(kT2TileBytes % (16 * kT2Pieces) == 0, "pieces of whole 16-byte units");

template <int X>
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,"
      "%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
        "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__global__ void __launch_bounds__(kT2Threads, 1) mel_to_linear_tc2_kernel(const TcGemmParams p, const int n_ftiles) {
  extern __shared__ __align__(1024) unsigned char tc_smem[];
  constexpr uint32_t kLbo = (128 / 8) * 128;       // byte stride between 8-element K blocks of a 128-row operand
  constexpr uint32_t kIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kT2Frames >> 3) << 17) | ((uint32_t)(kT2Bins >> 4) << 24);
  unsigned char* const smF = tc_smem;                            // frames: 3 parts
  unsigned char* const smP = tc_smem + kT2TileBytes;             // pseudo-inverse tiles: 2 buffers of 3 parts
  __shared__ __align__(8) uint64_t mbar[6];                      // 0,1 tile landed; 2,3 accumulator full; 4,5 accumulator drained
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (warp == 0) {
    const uint32_t dst = (uint32_t)__cvta_generic_to_shared(&tmem_base_s);
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst), "r"(256) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  const uint32_t bar0 = (uint32_t)__cvta_generic_to_shared(&mbar[0]);
  if (tid == 0) {
    for (int i = 0; i < 4; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0 + 8 * i) : "memory");
    for (int i = 4; i < 6; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar0 + 8 * i), "r"(kT2EpiWarps) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_s;
  const uint32_t smF_addr = (uint32_t)__cvta_generic_to_shared(smF);
  const uint32_t smP_addr = (uint32_t)__cvta_generic_to_shared(smP);
  const bool pow15 = p.mp.power == 1.5f;
  const bool producer = warp == kT2EpiWarps;
  uint32_t n_tile = 0;                             // (frame tile, bin tile) pairs this CTA has started: buffer = n_tile & 1, use = n_tile >> 1

  for (int ft = blockIdx.x; ft < n_ftiles; ft += gridDim.x) {
    const long long row0 = (long long)ft * kT2Frames;
    if (!producer) {
      // ---- frames: fp32 mel -> amplitude -> three bf16 terms, canonical K-major layout; thread (row, quarter) converts
      //      3 / 3 / 2 / 2 K blocks of 8.  Every MMA of the previous frame tile is complete: the epilogue warps have waited
      //      for its last accumulator.
      const int r = tid >> 2, quarter = tid & 3;
      const long long row = row0 + r;
      const int kb0 = quarter < 2 ? 3 * quarter : 6 + 2 * (quarter - 2), nkb = quarter < 2 ? 3 : 2;
      // all of the thread's 24 (16) values are requested before the first is used: one exposed load latency per frame tile
      float xin[24];
      const float* const arow = p.a + row * p.lda;
#pragma unroll
      for (int i = 0; i < 24; ++i) {
        const int k = kb0 * 8 + i;
        xin[i] = (i < 8 * nkb && row < p.mp.rows && k < p.k_total) ? __ldg(arow + k) : 0.0f;
      }
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        if (q < nkb) {
          __align__(16) __nv_bfloat16 h8[8], m8[8], l8[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int k = (kb0 + q) * 8 + i;
            const float x = (row < p.mp.rows && k < p.k_total) ? mel_in_value(xin[8 * q + i], p.in_kind, p.mp) : 0.0f;
            h8[i] = __float2bfloat16_rn(x);
            const float r1 = x - __bfloat162float(h8[i]);
            m8[i] = __float2bfloat16_rn(r1);
            l8[i] = __float2bfloat16_rn(r1 - __bfloat162float(m8[i]));
          }
          const uint32_t off = (uint32_t)(kb0 + q) * kLbo + (uint32_t)(r >> 3) * 128u + (uint32_t)(r & 7) * 16u;
          *reinterpret_cast<uint4*>(smF + off) = *reinterpret_cast<const uint4*>(h8);
          *reinterpret_cast<uint4*>(smF + kT2PartBytes + off) = *reinterpret_cast<const uint4*>(m8);
          *reinterpret_cast<uint4*>(smF + 2 * kT2PartBytes + off) = *reinterpret_cast<const uint4*>(l8);
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the tensor core
    }
    auto load_tile = [&](int bt, uint32_t nt) {                      // pseudo-inverse rows of bin tile bt -> buffer nt & 1
      const uint32_t b = nt & 1u;
      const unsigned char* src = reinterpret_cast<const unsigned char*>(p.b) + (size_t)bt * kT2TileBytes;
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar0 + 8 * b), "r"(kT2TileBytes) : "memory");
      // several copies in flight instead of one 60 KB copy (the copy engine works a single request off sequentially)
      constexpr uint32_t kPiece = kT2TileBytes / kT2Pieces;
#pragma unroll
      for (uint32_t i = 0; i < (uint32_t)kT2Pieces; ++i)
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smP_addr + b * kT2TileBytes + i * kPiece), "l"(src + i * kPiece), "r"(kPiece), "r"(bar0 + 8 * b) : "memory");
    };
    // buffer nt & 1 is free once the MMAs of pair nt - 2 are complete (accumulator-full barrier of that pair)
    auto wait_buffer_free = [&](uint32_t nt) {
      if (nt >= 2) mbar_wait(bar0 + 16 + 8 * (nt & 1u), ((nt - 2) >> 1) & 1u);
    };
    if (producer && lane == 0) {                                     // the first tile streams in behind the conversion
      wait_buffer_free(n_tile);
      load_tile(0, n_tile);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();                                                 // the frame operand is in place
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    if (producer) {
      if (lane == 0) {
        for (int bt = 0; bt < kT2Tiles; ++bt) {
          const uint32_t nt = n_tile + (uint32_t)bt, b = nt & 1u, use = nt >> 1;
          if (bt + 1 < kT2Tiles) { wait_buffer_free(nt + 1); load_tile(bt + 1, nt + 1); }
          mbar_wait(bar0 + 8 * b, use & 1u);                         // the tile has landed
          if (use >= 1) mbar_wait(bar0 + 32 + 8 * b, (use - 1) & 1u);  // the accumulator's previous contents are drained
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const int pa[6] = {0, 0, 1, 0, 1, 2};                      // (pinv part, frame part): hi*hi, hi*mid, mid*hi, hi*lo, mid*mid, lo*hi
          const int pb[6] = {0, 1, 0, 2, 1, 0};
          const uint32_t abase = smP_addr + b * kT2TileBytes;
          const uint32_t dcol = tmem_base + b * (uint32_t)kT2Frames;
#pragma unroll
          for (int c = 0; c < 6; ++c) {
#pragma unroll
            for (int j = 0; j < kTcKSteps; ++j) {
              const uint64_t da = umma_smem_desc(abase + pa[c] * kT2PartBytes + 2 * j * kLbo, kLbo, 128);
              const uint64_t db = umma_smem_desc(smF_addr + pb[c] * kT2PartBytes + 2 * j * kLbo, kLbo, 128);
              umma_bf16(dcol, da, db, kIdesc, (c | j) != 0 ? 1u : 0u);
            }
          }
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar0 + 16 + 8 * b) : "memory");
        }
      }
      __syncwarp();
    } else {
      // ---- epilogue warps: TMEM lane group = warp % 4 (bins), 32 frames = warp / 4
      const int lane_base = (warp & 3) * 32, fq = warp >> 2;
      // (kernel parameters read once: inside the unrolled store loop each use was a constant-bank load on the critical path)
      const int ldo = p.ldo, n_valid = p.n_valid;
      const long long n_rows = p.mp.rows;
      const bool do_power = p.out_kind == 1;
      const float power = p.mp.power;
      for (int bt = 0; bt < kT2Tiles; ++bt) {
        const uint32_t nt = n_tile + (uint32_t)bt, b = nt & 1u, use = nt >> 1;
        mbar_wait(bar0 + 16 + 8 * b, use & 1u);                      // the MMAs of this pair are complete
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        uint32_t v[32];
        tmem_ld32<0>(tmem_base + ((uint32_t)lane_base << 16) + b * (uint32_t)kT2Frames + 32u * (uint32_t)fq, v);
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar0 + 32 + 8 * b) : "memory");   // drained
        const int bin = bt * kT2Bins + lane_base + lane;
        if (bin < n_valid) {
          float* o = p.out + (row0 + 32 * fq) * ldo + bin;
          const int nrows = (int)min((long long)32, n_rows - (row0 + 32 * fq));
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float x = fmaxf(1e-10f, __uint_as_float(v[j]));
            if (do_power) {
              if (pow15) {                         // x ** 1.5 (the shipped `power`): x * sqrt(x), sqrt.approx is ~1 ulp
                float sq;
                asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(sq) : "f"(x));
                x *= sq;
              } else {
                x = exp2f(power * log2f(x));
              }
            }
            if (j < nrows) *o = x;
            o += ldo;
          }
        }
      }
    }
    n_tile += (uint32_t)kT2Tiles;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256) : "memory");
  }
}

}  // namespace ttsa
