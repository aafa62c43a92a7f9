// Shared definitions for the ttsa kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ttsa {

constexpr int kNfft = 2048;          // real transform length (num_freq 1025)
constexpr int kM = 1024;             // packed complex transform length
constexpr int kF = 1025;             // bins
constexpr int kNF = 8;               // frames per tile == warps per CTA (one warp owns one frame)
constexpr int kThreads = kNF * 32;
constexpr int kRowStride = 34;       // float2 row stride of the 32x32 exchange buffer (conflict-free 64b writes / 128b reads)
constexpr int kBufFloats = 32 * kRowStride * 2;   // per-warp exchange buffer, reused as the warp's overlap-add slot

// Geometry and scalar constants, passed by value to every frame kernel.
struct Geo {
  int hop, win, off0;        // off0 = n_fft/2 - (n_fft - win)/2: sample offset of window tap 0 relative to t*hop
  int half;                  // ceil(win / 2): number of packed complex inputs that are not identically zero
  int wlen;                  // half rounded up to 32: length of the even/odd window tables held in shared memory
  int carry_len;             // win - hop
  int span_len;              // (kNF-1)*hop + win
  int nwarm;                 // (win-1)/hop: earlier frames overlapping a segment's first owned sample
  int num_mels;
  float inv_hop;
  // shared memory layout (float offsets)
  int sm_plane0, sm_plane1, sm_carry0, sm_carry1, sm_wE, sm_wO, sm_pw, sm_tw, sm_g, sm_total;
  // spectrogram value -> magnitude:  S = exp2(c1 * clip(x, lo, hi) + c0)   (denormalize, +ref, db_to_amp, **power fused)
  float s_c1, s_c0, s_lo, s_hi;
  // amplitude -> normalised dB:      v = clip(n_a * log2(max(min_amp, a)) + n_b, n_lo, n_hi)
  float n_a, n_b, n_lo, n_hi, min_amp;
  float preemph;
};

// Plan-owned device tables.
struct Tables {
  const float2* tw;      // [32*32]  W_1024^(a*b) = (cos, -sin)(2 pi a b / 1024)
  const float2* g;       // [512]    G_k = -j * W_2048^k = (-sin, -cos)(pi k / 1024)
  const float* wE;       // [1024]   w[2q]   (zero padded)
  const float* wO;       // [1024]   w[2q+1]
  const float* pw;       // [hop]    1 / sum_q w[r + q*hop]^2   (interior window-sum-square, periodic in hop)
  // sparse mel basis (CSR over mel rows; each row is one contiguous run of bins)
  const int* mel_lo;     // [num_mels]
  const int* mel_cnt;    // [num_mels]
  const float* mel_val;  // [num_mels * mel_ld]
  int mel_ld;
};

// Batch layout on the device.
struct BatchDev {
  const int* T;              // [B] frames
  const int* wav_len;        // [B] samples
  const long long* frame_off;// [B+1]
  const long long* wav_off;  // [B+1] (multiples of 4)
  const int* tile_off;       // [B+1] prefix sum of ceil(T/kNF)
  int B;
  int total_tiles;
};

__device__ __forceinline__ int reflect_index(int i, int L) {
  // np.pad(..., mode='reflect') index map (triangle-wave fold); L >= 1
  if (i >= 0 && i < L) return i;
  if (L == 1) return 0;
  const int period = 2 * (L - 1);
  int m = i % period;
  if (m < 0) m += period;
  return m >= L ? period - m : m;
}

}  // namespace ttsa
