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
constexpr int kRowFloats = 68;       // row stride of the 32x32 exchange buffer: [re 0..31 | im 0..31 | pad], 68 = 4 mod 32
                                     // keeps both the 32-bit column writes and the 128-bit row reads conflict-free
constexpr int kBufFloats = 32 * kRowFloats;       // per-warp exchange buffer, reused as the warp's overlap-add slot
constexpr int kSlotPlane = 1040;     // slot = even-sample plane + odd-sample plane (16 banks apart)

// Geometry and scalar constants, passed by value to every frame kernel.
struct Geo {
  int hop, win, off0;        // off0 = n_fft/2 - (n_fft - win)/2: sample offset of window tap 0 relative to t*hop
  int half;                  // ceil(win / 2): number of packed complex inputs that are not identically zero
  int wlen;                  // floats in each paired window table held in shared memory (kernel class rows * 32)
  int plane_len;             // floats in each parity plane of the staged span (tail beyond the span stays zero)
  int carry_len;             // win - hop
  int span_len;              // (kNF-1)*hop + win
  int nwarm;                 // (win-1)/hop: earlier frames overlapping a segment's first owned sample
  int num_mels;
  float inv_hop;
  // shared memory layout (float offsets)
  int sm_plane0, sm_plane1, sm_carry0, sm_carry1, sm_wE, sm_wO, sm_pw, sm_wsyn, sm_tw, sm_g, sm_total;
  // spectrogram value -> magnitude:  S = exp2(c1 * clip(x, lo, hi) + c0)   (denormalize, +ref, db_to_amp, **power fused)
  float s_c1, s_c0, s_lo, s_hi;
  // amplitude -> normalised dB:      v = clip(n_a * log2(max(min_amp, a)) + n_b, n_lo, n_hi)
  float n_a, n_b, n_lo, n_hi, min_amp;
  float preemph;
};

// Plan-owned device tables.
struct Tables {
  const float4* tw4;     // [16][32]  (wr(2m), wr(2m+1), wi(2m), wi(2m+1)),  w(k) = W_1024^(lane*k) = (cos, -sin)(2 pi lane k / 1024)
  const float4* g4;      // [8][32]   (gx(2m), gx(2m+1), gy(2m), gy(2m+1)),  g(k1) = -j W_2048^(32 k1 + lane)
  const float2* wE2;     // [16][32]  (w[2q], w[2(q+32)]),  q = lane + 64 m   (even window taps, paired for two packed rows)
  const float2* wO2;     // [16][32]  (w[2q+1], w[2(q+32)+1])
  const float* wE;       // [1024]    w[2q]   (zero padded)
  const float* wO;       // [1024]    w[2q+1]
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
