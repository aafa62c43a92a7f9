// Small kernels around the frame kernels: elementwise API steps, layout shim, pre/de-emphasis, mel <-> linear.
#pragma once
#include "common.cuh"

namespace ttsa {

// ---------------------------------------------------------------------------------------------------------
// elementwise (_normalize / _denormalize / _amp_to_db / _db_to_amp, utils/audio.py:79-126)
// ---------------------------------------------------------------------------------------------------------
struct PwParams {
  int op;
  int signal_norm, symmetric_norm, clip_norm;
  float min_level_db, max_norm, min_amp;
};

__global__ void pointwise_kernel(PwParams p, const float* __restrict__ x, float* __restrict__ y, long long n) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    float v = x[i];
    if (p.op == 0) {            // _normalize
      if (p.signal_norm) {
        v = (v - p.min_level_db) / -p.min_level_db;
        if (p.symmetric_norm) {
          v = (2.0f * p.max_norm) * v - p.max_norm;
          if (p.clip_norm) v = fminf(fmaxf(v, -p.max_norm), p.max_norm);
        } else {
          v = p.max_norm * v;
          if (p.clip_norm) v = fminf(fmaxf(v, 0.0f), p.max_norm);
        }
      }
    } else if (p.op == 1) {     // _denormalize
      if (p.signal_norm) {
        if (p.symmetric_norm) {
          if (p.clip_norm) v = fminf(fmaxf(v, -p.max_norm), p.max_norm);
          v = ((v + p.max_norm) * -p.min_level_db / (2.0f * p.max_norm)) + p.min_level_db;
        } else {
          if (p.clip_norm) v = fminf(fmaxf(v, 0.0f), p.max_norm);
          v = (v * -p.min_level_db / p.max_norm) + p.min_level_db;
        }
      }
    } else if (p.op == 2) {     // _amp_to_db
      v = 20.0f * log10f(fmaxf(p.min_amp, v));
    } else {                    // _db_to_amp
      v = exp10f(v * 0.05f);
    }
    y[i] = v;
  }
}

// out[c, r] = in[r, c]
__global__ void transpose_kernel(const float* __restrict__ in, float* __restrict__ out, long long rows, long long cols) {
  __shared__ float tile[32][33];
  const long long c0 = (long long)blockIdx.x * 32, r0 = (long long)blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const long long r = r0 + j, c = c0 + threadIdx.x;
    if (r < rows && c < cols) tile[j][threadIdx.x] = in[r * cols + c];
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const long long c = c0 + j, r = r0 + threadIdx.x;
    if (r < rows && c < cols) out[c * rows + r] = tile[threadIdx.x][j];
  }
}

// ---------------------------------------------------------------------------------------------------------
// apply_preemphasis: y[n] = x[n] - p x[n-1]   (utils/audio.py:128-131)
// ---------------------------------------------------------------------------------------------------------
__global__ void preemphasis_kernel(BatchDev bd, float p, const float* __restrict__ x, float* __restrict__ y) {
  const int u = blockIdx.y;
  const int L = bd.wav_len[u];
  const float* xs = x + bd.wav_off[u];
  float* ys = y + bd.wav_off[u];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < L; i += gridDim.x * blockDim.x)
    ys[i] = fmaf(-p, i > 0 ? xs[i - 1] : 0.0f, xs[i]);
}

// ---------------------------------------------------------------------------------------------------------
// Waveform post-processing (the steps right after the synthesis path)
//   save_wav scaling      wav * (32767 / max(0.01, max|wav|)) -> int16      (utils/audio.py:56-58)
//   find_endpoint         first silent window                                (utils/audio.py:203-210)
//   sentence gaps         10 000 zero samples after every sentence           (server/synthesizer.py:157-161)
// Packed waveforms; an optional per-utterance length override (e.g. the endpoints) replaces BatchDev::wav_len.
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int wav_len_of(const BatchDev& bd, const int* lens, int u) {
  const int L = bd.wav_len[u];
  if (lens == nullptr) return L;
  const int v = lens[u];
  return v < 0 ? 0 : (v < L ? v : L);
}

// peaks[u] = max |wav_u|  (bit pattern of a non-negative float orders like an unsigned integer); peaks zeroed by the caller
__global__ void wav_peak_kernel(BatchDev bd, const int* __restrict__ lens, const float* __restrict__ wav, unsigned* __restrict__ peaks) {
  const int u = blockIdx.y;
  const int L = wav_len_of(bd, lens, u);
  const float* x = wav + bd.wav_off[u];
  float m = 0.0f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < L; i += gridDim.x * blockDim.x) m = fmaxf(m, fabsf(x[i]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m > 0.0f) atomicMax(peaks + u, __float_as_uint(m));
}

// endpoints[u] = first x + hop with max(wav[x : x + window]) < threshold, x = hop, 2 hop, ... < L - window; else L.
// One CTA per (candidate, utterance); endpoints pre-set to L by wav_endpoint_init_kernel.
__global__ void wav_endpoint_init_kernel(BatchDev bd, int* __restrict__ endpoints) {
  const int u = blockIdx.x * blockDim.x + threadIdx.x;
  if (u < bd.B) endpoints[u] = bd.wav_len[u];
}
__global__ void wav_endpoint_kernel(BatchDev bd, const float* __restrict__ wav, int window, int hop, double threshold,
                                    int* __restrict__ endpoints) {
  const int u = blockIdx.y;
  const int L = bd.wav_len[u];
  const int x0 = (blockIdx.x + 1) * hop;
  if (x0 >= L - window) return;
  const float* x = wav + bd.wav_off[u] + x0;
  float m = -INFINITY;                                     // the reference takes the SIGNED maximum, not max |x|
  for (int i = threadIdx.x; i < window; i += blockDim.x) m = fmaxf(m, x[i]);
  __shared__ float red[32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) m = fmaxf(m, red[w]);
    if ((double)m < threshold) atomicMin(endpoints + u, x0 + hop);
  }
}

// out_off[u] = sum_{v<u} (len_v + gap), out_off[B] = total; peaks -> per-utterance (or joint) scale
// 32767 / max(0.01, peak) in float64.  One CTA.
__global__ void pcm_plan_kernel(BatchDev bd, const int* __restrict__ lens, const unsigned* __restrict__ peaks, int joint,
                                long long gap, long long* __restrict__ out_off, double* __restrict__ scales) {
  __shared__ float s_peak;
  if (threadIdx.x == 0) {
    long long acc = 0;
    float pk = 0.0f;
    for (int u = 0; u < bd.B; ++u) {
      out_off[u] = acc;
      acc += wav_len_of(bd, lens, u) + gap;
      pk = fmaxf(pk, __uint_as_float(peaks[u]));
    }
    out_off[bd.B] = acc;
    s_peak = pk;
  }
  __syncthreads();
  for (int u = threadIdx.x; u < bd.B; u += blockDim.x) {
    const double pk = (double)(joint ? s_peak : __uint_as_float(peaks[u]));
    scales[u] = 32767.0 / fmax(0.01, pk);
  }
}

// int16 conversion: numpy astype(np.int16) of the scaled sample = truncation toward zero.  f32_arith: the product is
// formed in float32 with the float64 scale rounded to float32 (what the reference computes for a float32 waveform
// under its pinned numpy); otherwise in float64 (float64 waveform: the reference's default path, de-emphasis and the
// server's list concatenation both widen).
__global__ void pcm_convert_kernel(BatchDev bd, const int* __restrict__ lens, const float* __restrict__ wav,
                                   const long long* __restrict__ out_off, const double* __restrict__ scales, long long gap,
                                   int f32_arith, long long capacity, short* __restrict__ out) {
  const int u = blockIdx.y;
  const int L = wav_len_of(bd, lens, u);
  const float* x = wav + bd.wav_off[u];
  const long long o0 = out_off[u];
  const double sc = scales[u];
  const float scf = (float)sc;
  const long long n = L + gap;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    if (o0 + i >= capacity) break;
    short v = 0;
    if (i < L) v = f32_arith ? (short)(int)__fmul_rn(x[i], scf) : (short)(int)((double)x[i] * sc);
    out[o0 + i] = v;
  }
}

// ---------------------------------------------------------------------------------------------------------
// apply_inv_preemphasis: y[n] = x[n] + p y[n-1]   (utils/audio.py:133-136)
// first-order linear recurrence as a blocked scan: chunks of kDeChunk samples, one CTA each.
//   pass 1: zero-state response of every chunk at its last sample (chunk aggregate)
//   pass 2: carry-in = sum over earlier chunks of aggregate * (p^chunk)^(distance), then the in-chunk scan
// ---------------------------------------------------------------------------------------------------------
constexpr int kDeThreads = 256;
constexpr int kDePer = 8;
constexpr int kDeChunk = kDeThreads * kDePer;

struct DeParams {
  float p;
  float log2p;
  const int* chunk_off;   // [B+1] prefix sum of ceil(L/kDeChunk)
};

__device__ __forceinline__ float pow_p(float log2p, int k) { return exp2f(log2p * (float)k); }

// in-chunk scan with zero initial state; returns the thread's values and leaves the exclusive carry for this
// thread's first element in `carry_t`
__device__ __forceinline__ void de_local_scan(const float* xs, int base, int L, float p, float log2p,
                                              float (&v)[kDePer], float& carry_t, float* warp_sums) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float acc = 0.0f;
#pragma unroll
  for (int j = 0; j < kDePer; ++j) {
    const int i = base + tid * kDePer + j;
    const float xv = i < L ? xs[i] : 0.0f;
    acc = fmaf(p, acc, xv);
    v[j] = acc;
  }
  // inclusive scan of thread aggregates under  y_t = agg_t + p^kDePer * y_{t-1}
  float inc = acc;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const float o = __shfl_up_sync(0xffffffffu, inc, d);
    if (lane >= d) inc = fmaf(pow_p(log2p, kDePer * d), o, inc);
  }
  if (lane == 31) warp_sums[warp] = inc;
  __syncthreads();
  float wcarry = 0.0f;   // state at the end of the previous warp
  for (int w = 0; w < warp; ++w) wcarry = fmaf(pow_p(log2p, kDePer * 32), wcarry, warp_sums[w]);
  float excl = __shfl_up_sync(0xffffffffu, inc, 1);
  if (lane == 0) excl = 0.0f;
  // state just before this thread's first element
  carry_t = fmaf(pow_p(log2p, kDePer * lane), wcarry, excl);
}

__global__ void __launch_bounds__(kDeThreads)
deemph_aggregate_kernel(BatchDev bd, DeParams dp, const float* __restrict__ x, float* __restrict__ agg) {
  __shared__ float warp_sums[kDeThreads / 32];
  const int u = blockIdx.y;
  const int nchunks = dp.chunk_off[u + 1] - dp.chunk_off[u];
  const int L = bd.wav_len[u];
  const float* xs = x + bd.wav_off[u];
  for (int c = blockIdx.x; c < nchunks; c += gridDim.x) {
    float v[kDePer], carry_t;
    __syncthreads();
    de_local_scan(xs, c * kDeChunk, L, dp.p, dp.log2p, v, carry_t, warp_sums);
    if (threadIdx.x == kDeThreads - 1)
      agg[dp.chunk_off[u] + c] = fmaf(pow_p(dp.log2p, kDePer), carry_t, v[kDePer - 1]);
  }
}

__global__ void __launch_bounds__(kDeThreads)
deemph_apply_kernel(BatchDev bd, DeParams dp, const float* __restrict__ agg, const float* x, float* y) {
  __shared__ float warp_sums[kDeThreads / 32];
  __shared__ float s_in;
  const int u = blockIdx.y;
  const int nchunks = dp.chunk_off[u + 1] - dp.chunk_off[u];
  const int L = bd.wav_len[u];
  const float* xs = x + bd.wav_off[u];
  float* ys = y + bd.wav_off[u];
  const float pc = pow_p(dp.log2p, kDeChunk);
  for (int c = blockIdx.x; c < nchunks; c += gridDim.x) {
    __syncthreads();
    if (threadIdx.x == 0) {
      float st = 0.0f;   // y[c*chunk - 1]
      for (int k = 0; k < c; ++k) st = fmaf(pc, st, agg[dp.chunk_off[u] + k]);
      s_in = st;
    }
    float v[kDePer], carry_t;
    de_local_scan(xs, c * kDeChunk, L, dp.p, dp.log2p, v, carry_t, warp_sums);   // contains a __syncthreads
    // fold the chunk's carry-in: state before this thread = carry_t + p^(tid*kDePer) * s_in
    float st = fmaf(pow_p(dp.log2p, kDePer * threadIdx.x), s_in, carry_t);
    float pj = dp.p;
#pragma unroll
    for (int j = 0; j < kDePer; ++j) {
      const int i = c * kDeChunk + threadIdx.x * kDePer + j;
      if (i < L) ys[i] = fmaf(pj, st, v[j]);
      pj *= dp.p;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// mel <-> linear
// ---------------------------------------------------------------------------------------------------------
struct MelParams {
  // amplitude from a normalised dB value: exp2(a_c1 * clip(x, lo, hi) + a_c0)   (power NOT applied)
  float a_c1, a_c0, a_lo, a_hi;
  // normalised dB from amplitude (same constants as Geo)
  float n_a, n_b, n_lo, n_hi, min_amp;
  float power;
  int num_mels;
  int F;                 // bins per row (num_freq)
  long long rows;
};

__device__ __forceinline__ float mel_in_value(float x, int in_kind, const MelParams& p) {
  if (in_kind == 1) {
    x = fminf(fmaxf(x, p.a_lo), p.a_hi);
    return exp2f(fmaf(p.a_c1, x, p.a_c0));
  }
  return x;
}

// _linear_to_mel (utils/audio.py:60-62): the Slaney basis has <= ~55 contiguous non-zeros per row, so the product
// is a banded contraction.  One warp per frame: the converted row is staged in shared memory, mel row m is reduced
// by the lanes of the warp.
__global__ void __launch_bounds__(256)
linear_to_mel_kernel(MelParams p, Tables tb, const float* __restrict__ lin, float* __restrict__ mel, int in_kind,
                     int out_kind) {
  __shared__ float rowbuf[8][kF + 7];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * 8 + warp;
  if (row >= p.rows) return;
  const float* src = lin + row * kF;
  for (int k = lane; k < kF; k += 32) {
    float v = mel_in_value(src[k], in_kind, p);
    rowbuf[warp][k] = in_kind == 1 ? fabsf(v) : v;   // out_linear_to_mel takes np.abs (utils/audio.py:177)
  }
  __syncwarp();
  float* dst = mel + row * p.num_mels;
  for (int m = 0; m < p.num_mels; ++m) {
    const int lo = tb.mel_lo[m], cnt = tb.mel_cnt[m];
    const float* mv = tb.mel_val + m * tb.mel_ld;
    float acc = 0.0f;
    for (int c = lane; c < cnt; c += 32) acc = fmaf(__ldg(mv + c), rowbuf[warp][lo + c], acc);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) {
      float v = acc;
      if (out_kind == 2) {
        v = fmaf(p.n_a, log2f(fmaxf(p.min_amp, v)), p.n_b);
        v = fminf(fmaxf(v, p.n_lo), p.n_hi);
      }
      dst[m] = v;
    }
  }
}

// _mel_to_linear (utils/audio.py:64-66): lin[row, k] = max(1e-10, sum_m pinv[k, m] * amp[row, m]) (optionally ** power).
// fp32 SIMT GEMM, CTA tile 128 bins x 64 frames x K (num_mels <= 128), thread tile 4 bins x 8 frames.
constexpr int kMtlBins = 128, kMtlRows = 64, kMtlMaxK = 128;
__global__ void __launch_bounds__(256)
mel_to_linear_kernel(MelParams p, const float* __restrict__ pinvT /*[K][ldp]*/, int ldp,
                     const float* __restrict__ mel, float* __restrict__ lin, int in_kind, int out_kind) {
  extern __shared__ __align__(16) float sm[];
  float* Ps = sm;                          // [K][128]
  float* As = sm + p.num_mels * kMtlBins;  // [K][64]
  const int K = p.num_mels;
  const int bin0 = blockIdx.x * kMtlBins;
  const long long row0 = (long long)blockIdx.y * kMtlRows;
  for (int i = threadIdx.x; i < K * kMtlBins; i += 256) {
    const int m = i / kMtlBins, b = i % kMtlBins;
    Ps[i] = (bin0 + b) < p.F ? pinvT[(long long)m * ldp + bin0 + b] : 0.0f;
  }
  for (int i = threadIdx.x; i < K * kMtlRows; i += 256) {
    const int r = i / K, m = i % K;        // coalesced over m
    const long long row = row0 + r;
    As[m * kMtlRows + r] = row < p.rows ? mel_in_value(mel[row * K + m], in_kind, p) : 0.0f;
  }
  __syncthreads();
  const int tb_ = threadIdx.x & 31;        // bins  tb_*4 .. +3
  const int tr = threadIdx.x >> 5;         // rows  tr*8 .. +7
  float acc[8][4];
#pragma unroll
  for (int r = 0; r < 8; ++r)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[r][b] = 0.0f;
  for (int m = 0; m < K; ++m) {
    const float4 pv = *reinterpret_cast<const float4*>(Ps + m * kMtlBins + tb_ * 4);
    const float4 a0 = *reinterpret_cast<const float4*>(As + m * kMtlRows + tr * 8);
    const float4 a1 = *reinterpret_cast<const float4*>(As + m * kMtlRows + tr * 8 + 4);
    const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
    const float pb[4] = {pv.x, pv.y, pv.z, pv.w};
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int b = 0; b < 4; ++b) acc[r][b] = fmaf(av[r], pb[b], acc[r][b]);
  }
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    const long long row = row0 + tr * 8 + r;
    if (row >= p.rows) continue;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int k = bin0 + tb_ * 4 + b;
      if (k < p.F) {
        float v = fmaxf(1e-10f, acc[r][b]);
        if (out_kind == 1) v = powf(v, p.power);
        lin[row * p.F + k] = v;
      }
    }
  }
}

}  // namespace ttsa
