// Any-size path: STFT / iSTFT / Griffin-Lim for power-of-two transforms other than n_fft 2048
// (num_freq = n_fft/2 + 1 in {129 ... 2049}; every shipped config uses 1025 and runs the warp-per-frame kernels of
// frame_kernels.cuh instead).  Same reference semantics (utils/audio.py:138-201 with librosa's centred, reflect-padded,
// Hann-windowed STFT), simplest correct mapping: one CTA per frame, a complex radix-2 transform of the whole frame in
// shared memory (decimation in frequency forward, decimation in time inverse, so no bit-reversal pass is needed: the
// per-bin step works on the bit-reversed positions), overlap-add in ceil(win/hop) launch phases (frames t = p mod R of one
// phase do not overlap: plain read-modify-write, deterministic order) into a zeroed waveform followed by a
// window-sum-square normalisation pass.  Not tuned: it exists so that a config with another num_freq is served by
// the CUDA path instead of TTSA_ERR_UNSUPPORTED.
#pragma once
#include "aux_kernels.cuh"
#include "frame_kernels.cuh"

namespace ttsa {

struct GenGeo {
  int n_fft, logn, F, hop, win, lpad, off0, num_mels;
  float preemph;
  float s_c1, s_c0, s_lo, s_hi;              // as Geo
  float n_a, n_b, n_lo, n_hi, min_amp;
};

struct GenTables {
  const float2* tw;      // [n_fft/2]  exp(-2 pi j k / n_fft)
  const float* win;      // [win]      periodic Hann
  const int* mel_lo;
  const int* mel_cnt;
  const float* mel_val;
  int mel_ld;
};

constexpr int kGenThreads = 256;

__device__ __forceinline__ float2 cmulf(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }

// forward: natural order in, bit-reversed order out
__device__ __forceinline__ void gen_fft_dif(float2* buf, int logn, const float2* __restrict__ tw, int tid) {
  const int n = 1 << logn;
  for (int s = logn; s >= 1; --s) {
    const int half = 1 << (s - 1), step = n >> s;
    for (int i = tid; i < n / 2; i += kGenThreads) {
      const int grp = i >> (s - 1), j = i & (half - 1);
      const int ia = (grp << s) + j, ib = ia + half;
      const float2 x = buf[ia], y = buf[ib];
      buf[ia] = make_float2(x.x + y.x, x.y + y.y);
      buf[ib] = cmulf(make_float2(x.x - y.x, x.y - y.y), __ldg(tw + j * step));
    }
    __syncthreads();
  }
}

// inverse (unnormalised): bit-reversed order in, natural order out
__device__ __forceinline__ void gen_ifft_dit(float2* buf, int logn, const float2* __restrict__ tw, int tid) {
  const int n = 1 << logn;
  for (int s = 1; s <= logn; ++s) {
    const int half = 1 << (s - 1), step = n >> s;
    for (int i = tid; i < n / 2; i += kGenThreads) {
      const int grp = i >> (s - 1), j = i & (half - 1);
      const int ia = (grp << s) + j, ib = ia + half;
      float2 w = __ldg(tw + j * step);
      w.y = -w.y;
      const float2 x = buf[ia], y = cmulf(buf[ib], w);
      buf[ia] = make_float2(x.x + y.x, x.y + y.y);
      buf[ib] = make_float2(x.x - y.x, x.y - y.y);
    }
    __syncthreads();
  }
}

template <int SRC>
__device__ __forceinline__ float gen_spec_to_mag(float x, const GenGeo& g) {
  if constexpr (SRC == SRC_NORM_DB) {
    x = fminf(fmaxf(x, g.s_lo), g.s_hi);
    return exp2f(fmaf(g.s_c1, x, g.s_c0));
  } else {
    return fabsf(x);
  }
}

__device__ __forceinline__ float gen_amp_to_norm_db(float a, const GenGeo& g) {
  const float v = fmaf(g.n_a, log2f(fmaxf(g.min_amp, a)), g.n_b);
  return fminf(fmaxf(v, g.n_lo), g.n_hi);
}

// MODE / SRC as in frame_kernels.cuh.  Dynamic shared memory: n_fft float2 + (F + 1) floats.
template <int MODE, int SRC>
__global__ void __launch_bounds__(kGenThreads)
gen_frame_kernel(const GenGeo g, const GenTables tb, const BatchDev bd, const FrameArgs a) {
  extern __shared__ __align__(16) unsigned char gen_smem[];
  float2* buf = reinterpret_cast<float2*>(gen_smem);
  float* mag = reinterpret_cast<float*>(gen_smem + (size_t)g.n_fft * sizeof(float2));
  __shared__ float red[2][kGenThreads / 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int N = g.n_fft, logn = g.logn, F = g.F;
  const float inv_n = 1.0f / (float)N;
  const int rsh = 32 - logn;

  for (long long row = blockIdx.x; row < a.rows_total; row += gridDim.x) {
    int u = 0;
    {
      int lo = 0, hi = bd.B;                 // largest u with frame_off[u] <= row
      while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (bd.frame_off[mid] <= row) lo = mid; else hi = mid;
      }
      u = lo;
    }
    const int t = (int)(row - bd.frame_off[u]);
    if (t >= bd.T[u]) continue;              // padding row of a strided layout
    if (MODE != MODE_ANALYSIS && a.ola_phases > 1 && (t % a.ola_phases) != a.ola_phase) continue;   // another launch's frame
    const int L = bd.wav_len[u];
    const long long woff = bd.wav_off[u];
    __syncthreads();                         // the previous row is done with buf / mag

    if constexpr (MODE != MODE_SYNTH) {
      // ---- frame: reflect-padded signal x periodic Hann centred in n_fft (librosa.stft, utils/audio.py:191-197)
      const float* src = a.wav_in + woff;
      for (int n = tid; n < N; n += kGenThreads) {
        const int m = n - g.lpad;
        float v = 0.0f;
        if (m >= 0 && m < g.win) {
          const int j = reflect_index(t * g.hop - N / 2 + n, L);
          float x = src[j];
          if (MODE == MODE_GL_ITER && a.wav_prev != nullptr) x = fmaf(-a.beta, a.wav_prev[woff + j], x);   // momentum mode
          if (MODE == MODE_ANALYSIS && a.preemph) x = fmaf(-g.preemph, j > 0 ? src[j - 1] : 0.0f, x);
          v = x * tb.win[m];
        }
        buf[n] = make_float2(v, 0.0f);
      }
      __syncthreads();
      gen_fft_dif(buf, logn, tb.tw, tid);    // X[k] at buf[brev(k)]
    }

    if constexpr (MODE == MODE_ANALYSIS) {
      if constexpr (SRC == OUT_COMPLEX) {
        float2* out = reinterpret_cast<float2*>(a.cplx_out) + row * F;
        for (int k = tid; k < F; k += kGenThreads) out[k] = buf[__brev((unsigned)k) >> rsh];
      } else {
        for (int k = tid; k < F; k += kGenThreads) {
          const float2 X = buf[__brev((unsigned)k) >> rsh];
          mag[k] = sqrtf(X.x * X.x + X.y * X.y);
        }
        __syncthreads();
        if (a.lin_out != nullptr) {
          float* out = a.lin_out + row * F;
          for (int k = tid; k < F; k += kGenThreads) out[k] = gen_amp_to_norm_db(mag[k], g);
        }
        if (a.mel_out != nullptr) {
          float* out = a.mel_out + row * g.num_mels;
          for (int m = warp; m < g.num_mels; m += kGenThreads / 32) {
            const int lo = tb.mel_lo[m], cnt = tb.mel_cnt[m];
            const float* mv = tb.mel_val + (size_t)m * tb.mel_ld;
            float acc = 0.0f;
            for (int c = lane; c < cnt; c += 32) acc = fmaf(mv[c], mag[lo + c], acc);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            if (lane == 0) out[m] = gen_amp_to_norm_db(acc, g);
          }
        }
      }
    } else {
      // ---- spectrum of the frame to synthesise, written hermitian at the bit-reversed positions
      float sc_num = 0.0f, sc_den = 0.0f;
      for (int k = tid; k < F; k += kGenThreads) {
        float2 Y;
        if constexpr (MODE == MODE_GL_ITER) {
          const float S = gen_spec_to_mag<SRC>(a.spec[row * F + k], g);
          const float2 X = buf[__brev((unsigned)k) >> rsh];
          const float m2 = X.x * X.x + X.y * X.y;
          if (m2 > 1e-37f) {                           // np.angle(0) = 0
            const float f = S * rsqrtf(m2);
            Y = make_float2(X.x * f, X.y * f);
          } else {
            Y = make_float2(S, 0.0f);
          }
          const float d = sqrtf(m2) - S;
          sc_num += d * d;
          sc_den += S * S;
        } else if constexpr (SRC == SRC_COMPLEX) {
          Y = reinterpret_cast<const float2*>(a.cplx_in)[row * F + k];
        } else {
          const float S = gen_spec_to_mag<SRC>(a.spec[row * F + k], g);
          float sn, cs;
          if (a.angles != nullptr) sincosf(a.angles[row * F + k], &sn, &cs);
          else sincospif(2.0f * philox_uniform4(a.seed, (unsigned long long)row * (unsigned long long)F + (unsigned long long)k).x, &sn, &cs);
          Y = make_float2(S * cs, S * sn);
        }
        if (k == 0 || k == N / 2) {
          buf[__brev((unsigned)k) >> rsh] = make_float2(Y.x, 0.0f);          // irfft ignores Im of DC / Nyquist
        } else {
          buf[__brev((unsigned)k) >> rsh] = Y;
          buf[__brev((unsigned)(N - k)) >> rsh] = make_float2(Y.x, -Y.y);
        }
      }
      if (MODE == MODE_GL_ITER && a.sc_acc != nullptr) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          sc_num += __shfl_xor_sync(0xffffffffu, sc_num, o);
          sc_den += __shfl_xor_sync(0xffffffffu, sc_den, o);
        }
        if (lane == 0) { red[0][warp] = sc_num; red[1][warp] = sc_den; }
      }
      __syncthreads();
      if (MODE == MODE_GL_ITER && a.sc_acc != nullptr && tid == 0) {
        float sn = 0.0f, sd = 0.0f;
        for (int w = 0; w < kGenThreads / 32; ++w) { sn += red[0][w]; sd += red[1][w]; }
        atomicAdd(a.sc_acc + 2 * u, sn);
        atomicAdd(a.sc_acc + 2 * u + 1, sd);
      }
      gen_ifft_dit(buf, logn, tb.tw, tid);
      // ---- window and overlap-add (librosa.istft, utils/audio.py:199-201); normalised by gen_wss_kernel afterwards
      float* dst = a.wav_out + woff;
      for (int m = tid; m < g.win; m += kGenThreads) {
        const int i = t * g.hop - g.off0 + m;
        if (i >= 0 && i < L) dst[i] += buf[g.lpad + m].x * inv_n * tb.win[m];   // frames of one launch phase never overlap
      }
    }
  }
}

// y[i] /= sum_t w[i + off0 - t hop]^2 over the frames that exist, where that sum exceeds tiny (librosa.istft)
__global__ void gen_wss_kernel(const GenGeo g, const GenTables tb, const BatchDev bd, float* __restrict__ wav) {
  const int u = blockIdx.y;
  const int L = bd.wav_len[u], T = bd.T[u];
  float* y = wav + bd.wav_off[u];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < L; i += gridDim.x * blockDim.x) {
    const int p = i + g.off0;
    int t_hi = p / g.hop;
    if (t_hi > T - 1) t_hi = T - 1;
    float ws = 0.0f;
    for (int t = t_hi; t >= 0; --t) {
      const int m = p - t * g.hop;
      if (m >= g.win) break;
      const float w = tb.win[m];
      ws = fmaf(w, w, ws);
    }
    if (ws > 1.17549435e-38f) y[i] = y[i] / ws;
  }
}

// linear -> mel for any F: one warp per row straight from global memory
__global__ void __launch_bounds__(256)
gen_linear_to_mel_kernel(MelParams p, int F, GenTables tb, const float* __restrict__ lin, float* __restrict__ mel, int in_kind,
                         int out_kind) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * 8 + warp;
  if (row >= p.rows) return;
  const float* src = lin + row * F;
  float* dst = mel + row * p.num_mels;
  for (int m = 0; m < p.num_mels; ++m) {
    const int lo = tb.mel_lo[m], cnt = tb.mel_cnt[m];
    const float* mv = tb.mel_val + (size_t)m * tb.mel_ld;
    float acc = 0.0f;
    for (int c = lane; c < cnt; c += 32) {
      const float v = mel_in_value(src[lo + c], in_kind, p);
      acc = fmaf(mv[c], in_kind == 1 ? fabsf(v) : v, acc);
    }
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

}  // namespace ttsa
