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

// compile-time integer usable as a generic-lambda argument in device code
template <int V>
struct IntC {
  static constexpr int value = V;
  __host__ __device__ constexpr operator int() const { return V; }
};

// Shared-memory layout and derived lengths of a (hop, win) geometry in kernel class `nz` (float offsets).  One
// constexpr function serves the host (plan creation) and the kernels: the shipped geometries are instantiated with
// compile-time (hop, win), so every offset, bound and stride below folds into an immediate; any other geometry
// runs the same code with the values read from Geo.
struct Layout {
  int hop, win, off0, half, wlen, plane_len, carry_len, span_len, nwarm;
  int sm_plane0, sm_plane1, sm_carry0, sm_wE, sm_wO, sm_pw, sm_wsyn, sm_tw, sm_g, sm_mbar, sm_total;
  int image_floats;          // [sm_wE, sm_mbar): the constant tables, one contiguous image built by the host
};
__host__ __device__ constexpr int round_up_c(int x, int m) { return (x + m - 1) / m * m; }
__host__ __device__ constexpr Layout make_layout(int hop, int win, int nz) {
  Layout l{};
  l.hop = hop;
  l.win = win;
  l.off0 = kNfft / 2 - (kNfft - win) / 2;
  l.half = (win + 1) / 2;
  l.wlen = nz * 32;
  l.carry_len = win - hop;
  l.span_len = (kNF - 1) * hop + win;
  l.nwarm = (win - 1) / hop;
  int off = kNF * kBufFloats;
  // a frame's loads start at up to ((kNF-1)*hop + 1)/2 and reach nz*32 packed samples further (zero tail)
  l.plane_len = round_up_c(((kNF - 1) * hop + 1) / 2 + 1 + nz * 32, 32);
  l.sm_plane0 = off; off += l.plane_len + 16;     // plane1 starts 16 banks away from plane0
  l.sm_plane1 = off; off += l.plane_len;
  l.sm_carry0 = off; off += round_up_c(l.carry_len + 1, 4);
  l.sm_wE = off; off += l.wlen;
  l.sm_wO = off; off += l.wlen;
  l.sm_pw = off; off += round_up_c(hop, 4);
  l.sm_wsyn = off; off += round_up_c(win, 4);
  l.sm_tw = off; off += 2048;
  l.sm_g = off; off += 1024;
  l.image_floats = off - l.sm_wE;                 // a multiple of 4 floats starting on a 16-byte boundary
  l.sm_mbar = off; off += 4;                      // mbarrier of the table copy
  l.sm_total = off;
  return l;
}

// Geometry and scalar constants, passed by value to every frame kernel.
struct Geo {
  Layout ly;
  int num_mels;
  int mel_smem_floats;       // > 0: the compact mel basis (Tables::mel_compact) fits behind Layout::sm_total in the
                             //      feature kernel's shared memory; 0: read the banded basis from global memory
  int mel_steps;             // > 0: steps of the lane schedule (Tables::mel_sched) of the warp-stream feature kernel
  int mel_seg_pairs[3];      // [0] > 0: the segment schedule (Tables::mel_seg) serves the warp-stream feature kernel instead:
                             //      step pairs of its three slots (host_tables.hpp, mel_segment_schedule)
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
  const float2* pw2;     // [hop]    (pw[r], pw[(r + 1) % hop]) / n_fft   (sample pairs of the warp-stream Griffin-Lim kernel)
  const float* edge_head;   // [max(0, warm*hop - win/2)]  1 / (n_fft wss) of an utterance's first samples (frames before 0 missing)
  const float* edge_tail;   // [max(0, win/2 - hop)]       the same for its last samples (frame T missing); gl_stream.cuh
  const float* wps_image;   // table image of the warp-stream Griffin-Lim kernel (gl_stream.cuh, WpsGeo): tw4 | g4 | wA | wB | pwx
  const float* smem_image;  // [Layout::image_floats]  wE2 | wO2 | pw / n_fft | signed synthesis window | tw4 | g4, laid out
                            // exactly as the kernels keep them in shared memory (one bulk copy per CTA)
  // sparse mel basis (CSR over mel rows; each row is one contiguous run of bins)
  const int* mel_lo;     // [num_mels]
  const int* mel_cnt;    // [num_mels]
  const float* mel_val;  // [num_mels * mel_ld]
  int mel_ld;
  // the same basis without padding: int2 (first tap index in mel_cval, first bin) per filter, one extra entry, then the taps
  const float* mel_compact;
  // lane schedule of the same basis for feat_stream.cuh: float4 (w0, w1, w2, bin) [Geo::mel_steps][32], then int [3][32] filter ids
  const float* mel_sched;
  // segment schedule (host_tables.hpp, MelSegSchedule::words): float4 [NP][32] | u32 [NP][32] | u32 [3][32]
  const unsigned* mel_seg;
};

// Batch layout on the device.
struct BatchDev {
  const int* T;              // [B] frames
  const int* wav_len;        // [B] samples
  const long long* frame_off;// [B+1]
  const long long* wav_off;  // [B+1] (multiples of 4)
  const int* tile_off;       // [B+1] prefix sum of ceil(T/kNF)
  const int* fine_off;       // [B+1] prefix sum of the fine segments per utterance (frame_kernel<..., FINE>)
  const int* tsum;           // [B+1] prefix sum of the frame counts (flattened frame index; feat_stream.cuh), or nullptr when it overflows int
  int B;
  int total_tiles;
  int total_fine;
};

#ifndef TTSA_WPS_WARPS
#define TTSA_WPS_WARPS 16
#endif
constexpr int kWpsWinStride = 20;           // floats per lane row of the warp-stream kernel's window tables (>= rows of a frame; stride / 4 odd)
constexpr int kWpsWarps = TTSA_WPS_WARPS;   // warps per CTA (one CTA per SM) of the warp-stream Griffin-Lim kernel
constexpr int kMelSegScratch = 196;         // floats per warp: 192 partial sums of the segment schedule, the zero cell, padding

// Warp index as a value the compiler knows to be the same in all 32 lanes (a shuffle from lane 0): what derives from it
// (the warp's run of frames, utterance bounds, buffer addresses, loop counters) can then live in uniform registers
// instead of the 128 per-thread ones.  -DTTSA_WARP_UNIFORM=0 restores the plain tid >> 5, =1 broadcasts only the warp index.
#ifndef TTSA_WARP_UNIFORM
#define TTSA_WARP_UNIFORM 2
#endif
__device__ __forceinline__ int warp_index() {
#if TTSA_WARP_UNIFORM
  return __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
#else
  return (int)(threadIdx.x >> 5);
#endif
}
// The same for a value every lane loaded from the same address (per-run scalars of the batch descriptor).
__device__ __forceinline__ int uni(int v) {
#if TTSA_WARP_UNIFORM >= 2
  return __shfl_sync(0xffffffffu, v, 0);
#else
  return v;
#endif
}
__device__ __forceinline__ long long uni(long long v) {
#if TTSA_WARP_UNIFORM >= 2
  const int lo = __shfl_sync(0xffffffffu, (int)(unsigned)(unsigned long long)v, 0);
  const int hi = __shfl_sync(0xffffffffu, (int)((unsigned long long)v >> 32), 0);
  return (long long)(((unsigned long long)(unsigned)hi << 32) | (unsigned)lo);
#else
  return v;
#endif
}
// floats of the mel schedule (or compact basis) the warp-stream feature kernel keeps in shared memory
__host__ __device__ inline int feat_mel_floats(const Geo& g) {
  if (g.mel_seg_pairs[0] > 0) return 160 * (g.mel_seg_pairs[0] + g.mel_seg_pairs[1] + g.mel_seg_pairs[2]) + 96 + kWpsWarps * kMelSegScratch;
  return g.mel_steps > 0 ? g.mel_steps * 128 + 96 : g.mel_smem_floats;
}

// Work partition of the warp-stream Griffin-Lim kernel (gl_stream.cuh), built by the host with the batch.
struct WpsDev {
  const int* cut;     // [grid * kWpsWarps + 1] flattened frame index where each warp's range starts (non-decreasing)
  const int* tsum;    // [B + 1] prefix sum of the frame counts (flattened frame index of each utterance's frame 0)
  const int* u0;      // [grid * kWpsWarps] utterance that holds the first frame of each warp's range
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
