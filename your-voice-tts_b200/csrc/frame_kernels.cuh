// Frame kernels: one warp owns one STFT frame (n_fft 2048 as a packed complex 1024-point FFT = 32 x 32,
// one shared-memory exchange per transform), one CTA streams over runs of consecutive frames of an utterance.
//
//   MODE_GL_ITER   y -> [window, FFT] -> |S| e^{j angle X} -> [iFFT, window, overlap-add, / wss] -> y'
//                  (utils/audio.py:186-188, one Griffin-Lim iteration in ONE HBM round trip)
//   MODE_SYNTH     spectrum (|S| e^{j theta0} or complex) -> iFFT -> overlap-add        (utils/audio.py:183-185, 199-201)
//   MODE_ANALYSIS  wav -> [pre-emphasis, reflect pad, window, FFT] -> complex or normalised dB linear/mel
//                  (utils/audio.py:138-152, 191-197)
//
// The window (win taps, centred in n_fft) is shorter than the transform, and a circular shift of the frame is a
// pure phase ramp that the per-bin projection S*X/|X| and the inverse transform undo exactly, so every frame is
// processed in "shifted" coordinates: tap m of the window is sample m of the transform, only ceil(win/2) packed
// inputs are non-zero and only the first win outputs of the inverse are needed.  (MODE_ANALYSIS with complex
// output and MODE_SYNTH with complex input or injected phases apply the ramp explicitly.)
//
// The inverse transform runs through the SAME forward code: ifft(Z) = conj(fft(conj(Z))); the per-bin step emits
// conj(Z') and the final window multiply negates the imaginary parts.  The transform body therefore exists once in
// the instruction stream (a rolled loop of four 32-point passes), which keeps the kernel inside the instruction cache.
#pragma once
#include "common.cuh"
#include "fft32.cuh"

namespace ttsa {

enum { MODE_GL_ITER = 0, MODE_SYNTH = 1, MODE_ANALYSIS = 2 };
enum { SRC_MAG = 0, SRC_NORM_DB = 1, SRC_COMPLEX = 2 };   // GL_ITER / SYNTH input kind
enum { OUT_COMPLEX = 0, OUT_FEATURES = 1 };               // ANALYSIS output kind

struct FrameArgs {
  const float* spec;      // [sum_T, F]      GL_ITER / SYNTH (mag or normalised dB)
  const float* spec_end;  // one past the last element of spec (bounds for the 16-byte row copies)
  const float* angles;    // [sum_T, F]      SYNTH: initial phases in radians, or nullptr -> counter RNG
  const float* cplx_in;   // [sum_T, F, 2]   SYNTH with SRC_COMPLEX
  const float* wav_in;    // packed wav      GL_ITER (previous y) / ANALYSIS
  float* wav_out;         // packed wav      GL_ITER / SYNTH
  float* cplx_out;        // [sum_T, F, 2]   ANALYSIS OUT_COMPLEX
  float* lin_out;         // [sum_T, F]      ANALYSIS OUT_FEATURES (nullable)
  float* mel_out;         // [sum_T, mels]   ANALYSIS OUT_FEATURES (nullable)
  float* sc_acc;          // [B, 2]          GL_ITER with SC: (sum (|X|-S)^2, sum S^2)
  unsigned long long seed;
  int preemph;            // ANALYSIS: apply y[n] - p*y[n-1] while staging
};

// ---------------------------------------------------------------------------------------------------------
// counter-based RNG (Philox4x32-10) for the initial phases when none are injected
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float philox_uniform(unsigned long long seed, unsigned long long ctr) {
  unsigned int c0 = (unsigned int)ctr, c1 = (unsigned int)(ctr >> 32), c2 = 0x243F6A88u, c3 = 0x85A308D3u;
  unsigned int k0 = (unsigned int)seed, k1 = (unsigned int)(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const unsigned int hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const unsigned int hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const unsigned int n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return (float)(c0 >> 8) * (1.0f / 16777216.0f);   // [0, 1)
}

// ---------------------------------------------------------------------------------------------------------
// asynchronous global -> shared copies
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gsrc) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// Asynchronous copy of one spectrogram row (kF floats, 4-byte aligned) into the warp's exchange buffer with 16-byte
// chunks taken from the enclosing 16-byte-aligned range; element k lands at dst[off + k], off = returned value.
// Chunks that would touch memory outside [lo, hi) fall back to 4-byte copies of the in-range elements.
__device__ __forceinline__ int row_to_smem_async(float* dst, const float* row_ptr, const float* lo, const float* hi, int lane) {
  const int off = (int)((reinterpret_cast<unsigned long long>(row_ptr) >> 2) & 3ull);
  const float* base = row_ptr - off;                      // 16-byte aligned
  const int nchunks = (off + kF + 3) >> 2;
  for (int c = lane; c < nchunks; c += 32) {
    const float* gsrc = base + 4 * c;
    if (gsrc >= lo && gsrc + 4 <= hi) {
      cp_async16(dst + 4 * c, gsrc);
    } else {
#pragma unroll
      for (int e = 0; e < 4; ++e)
        if (gsrc + e >= lo && gsrc + e < hi) cp_async4(dst + 4 * c + e, gsrc + e);
    }
  }
  return off;
}

// magnitude of one spectrogram value
template <int SRC>
__device__ __forceinline__ float spec_to_mag(float x, const Geo& g) {
  if constexpr (SRC == SRC_NORM_DB) {
    x = fminf(fmaxf(x, g.s_lo), g.s_hi);
    return exp2f(fmaf(g.s_c1, x, g.s_c0));
  } else {
    return fabsf(x);
  }
}

__device__ __forceinline__ float amp_to_norm_db(float a, const Geo& g) {
  const float v = fmaf(g.n_a, log2f(fmaxf(g.min_amp, a)), g.n_b);
  return fminf(fmaxf(v, g.n_lo), g.n_hi);
}

// unit phasor of the circular shift between window-relative and frame-relative coordinates, bin k
__device__ __forceinline__ float2 shift_phasor(int k, int lpad, float sign) {
  float sn, cs;
  sincospif(sign * (float)((k * lpad) & (kNfft - 1)) * (2.0f / (float)kNfft), &sn, &cs);
  return make_float2(cs, sn);
}
__device__ __forceinline__ float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }

// ---------------------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------------------
template <int MODE, int SRC, int NZ, bool SC>
__global__ void __launch_bounds__(kThreads, 2)
frame_kernel(const Geo g, const Tables tb, const BatchDev bd, const FrameArgs a) {
  extern __shared__ __align__(16) float smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr float kInvN = 1.0f / (float)kNfft;
  // kernel class: NZ 20 = "standard" geometry (<= 5 window taps per residue mod hop, 5*hop floats fit a slot, so slots
  // are zero padded to 5*hop and the overlap-add needs no tap predicates); NZ 32 = anything up to win <= 9*hop.
  constexpr bool kStd = (NZ <= 20);
  constexpr int ND = kStd ? 5 : kNF + 1;
  constexpr int kStage = 13;                      // span samples per thread prefetched through registers (13*256 >= 7*hop+win
                                                  // for every shipped geometry; any remainder is copied synchronously)

  float2* const buf = reinterpret_cast<float2*>(smem + warp * kBufFloats);
  float* const plane0 = smem + g.sm_plane0;
  float* const plane1 = smem + g.sm_plane1;
  float* const carry = smem + g.sm_carry0;
  float* const wE = smem + g.sm_wE;
  float* const wO = smem + g.sm_wO;
  float* const pw = smem + g.sm_pw;
  float2* const tw = reinterpret_cast<float2*>(smem + g.sm_tw);
  float2* const gt = reinterpret_cast<float2*>(smem + g.sm_g);

  // plan tables -> shared memory (once per persistent CTA)
  for (int i = tid; i < 1024; i += kThreads) tw[i] = tb.tw[i];
  for (int i = tid; i < 512; i += kThreads) gt[i] = tb.g[i];
  for (int i = tid; i < g.wlen; i += kThreads) { wE[i] = tb.wE[i]; wO[i] = tb.wO[i]; }
  if constexpr (MODE != MODE_ANALYSIS)
    for (int i = tid; i < g.hop; i += kThreads) pw[i] = tb.pw[i] * kInvN;     // 1/wss and the 1/n_fft of the inverse FFT

  // contiguous tile range of this CTA over the flattened (utterance, tile) list
  const long long tile_lo = (long long)blockIdx.x * bd.total_tiles / gridDim.x;
  const long long tile_hi = (long long)(blockIdx.x + 1) * bd.total_tiles / gridDim.x;
  int u = 0;
  {
    int lo = 0, hi = bd.B;   // largest u with tile_off[u] <= tile_lo
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (bd.tile_off[mid] <= tile_lo) lo = mid; else hi = mid;
    }
    u = lo;
  }

  long long tile = tile_lo;
  while (tile < tile_hi) {
    while (u + 1 < bd.B && tile >= bd.tile_off[u + 1]) ++u;
    const int T = bd.T[u];
    const int toff = bd.tile_off[u];
    const int ja = (int)(tile - toff);
    const long long seg_end = tile_hi < (long long)bd.tile_off[u + 1] ? tile_hi : (long long)bd.tile_off[u + 1];
    const int jb = (int)(seg_end - toff);
    tile = seg_end;
    const int L = bd.wav_len[u];                 // signal length (GL / SYNTH: hop*(T-1))
    if (MODE != MODE_ANALYSIS && L <= 0) continue;
    const long long frow0 = bd.frame_off[u];
    const long long woff = bd.wav_off[u];
    const float* __restrict__ src = (MODE != MODE_SYNTH) ? a.wav_in + woff : nullptr;

    // frames of earlier tiles still overlap this segment's first owned sample: recompute them (no output)
    const bool warm = (MODE != MODE_ANALYSIS) && ja > 0 && g.nwarm > 0;
    const int first_needed = warm ? ja * kNF - g.nwarm : 0;
    const int jt_first = warm ? ja - 1 : ja;
    bool has_carry = false;
    float sc_num = 0.0f, sc_den = 0.0f;

    // ---- span staging helpers: global -> registers -> (later) shared, de-interleaved by parity so that the
    //      re/im frame loads are conflict-free
    float stg[kStage];
    auto stage_load = [&](int jt) {
      const int i0 = jt * kNF * g.hop - g.off0;
      if (i0 >= 1 && i0 + g.span_len <= L) {                       // interior span: no reflection
        const float* __restrict__ sp = src + i0;
#pragma unroll
        for (int e = 0; e < kStage; ++e) {
          const int s = tid + e * kThreads;
          float val = 0.0f;
          if (s < g.span_len) {
            val = __ldg(sp + s);
            if (MODE == MODE_ANALYSIS && a.preemph) val = fmaf(-g.preemph, __ldg(sp + s - 1), val);
          }
          stg[e] = val;
        }
      } else {
#pragma unroll
        for (int e = 0; e < kStage; ++e) {
          const int s = tid + e * kThreads;
          float val = 0.0f;
          if (s < g.span_len) {
            const int j = reflect_index(i0 + s, L);
            val = __ldg(src + j);
            if (MODE == MODE_ANALYSIS && a.preemph) val = fmaf(-g.preemph, j > 0 ? __ldg(src + j - 1) : 0.0f, val);
          }
          stg[e] = val;
        }
      }
    };
    auto stage_store = [&](int jt) {
#pragma unroll
      for (int e = 0; e < kStage; ++e) {
        const int s = tid + e * kThreads;
        if (s < g.span_len) ((s & 1) ? plane1 : plane0)[s >> 1] = stg[e];
      }
      const int i0 = jt * kNF * g.hop - g.off0;
      for (int s = tid + kStage * kThreads; s < g.span_len; s += kThreads) {     // spans longer than the register window
        const int j = reflect_index(i0 + s, L);
        float val = __ldg(src + j);
        if (MODE == MODE_ANALYSIS && a.preemph) val = fmaf(-g.preemph, j > 0 ? __ldg(src + j - 1) : 0.0f, val);
        ((s & 1) ? plane1 : plane0)[s >> 1] = val;
      }
    };

    __syncthreads();                               // previous segment is done with planes / slots / carry
    if constexpr (MODE != MODE_SYNTH) {
      stage_load(jt_first);
      stage_store(jt_first);
      __syncthreads();
    }

    for (int jt = jt_first; jt < jb; ++jt) {
      const int t0 = jt * kNF;
      const int i0 = t0 * g.hop - g.off0;        // sample index of span position 0
      const bool write_out = jt >= ja;

      const int t = t0 + warp;
      if (t < T && t >= first_needed) {
        const long long row = frow0 + t;
        const bool own = t >= ja * kNF;            // warm-up frames are copies of another segment's frames
        float2 v[32];
        int s_off = 0;
        constexpr int kHalfBegin = (MODE == MODE_SYNTH) ? 1 : 0;
        constexpr int kHalfEnd = (MODE == MODE_ANALYSIS) ? 1 : 2;

#pragma unroll 1
        for (int half = kHalfBegin; half < kHalfEnd; ++half) {
          if (half == 0) {
            // ---------------------------------------------------------------- window the frame
            if constexpr (MODE != MODE_SYNTH) {
              const int o = warp * g.hop;
              const float* re_p = (o & 1) ? plane1 + ((o - 1) >> 1) : plane0 + (o >> 1);
              const float* im_p = (o & 1) ? plane0 + ((o + 1) >> 1) : plane1 + (o >> 1);
#pragma unroll
              for (int n2 = 0; n2 < 32; ++n2) {
                if (n2 < NZ) {
                  const int q = lane + 32 * n2;
                  const bool in = q < g.half;
                  v[n2].x = in ? re_p[q] * wE[q] : 0.0f;
                  v[n2].y = in ? im_p[q] * wO[q] : 0.0f;
                } else {
                  v[n2] = make_float2(0.0f, 0.0f);
                }
              }
            }
          } else if constexpr (MODE != MODE_ANALYSIS) {
            // ---------------------------------------------------------------- per-bin step -> conj(Z')
            // lane holds Z[32 k1 + lane]; bin k pairs with 1024 - k, held by lane (32 - lane) & 31.  Each lane
            // processes the 16 pairs whose first member is its own register k1 < 16.
            const int partner = (32 - lane) & 31;
            const bool l0 = lane == 0;
            const int lpad = (kNfft - g.win) >> 1;
            float2 r[16];
            if constexpr (MODE == MODE_GL_ITER) {
#pragma unroll
              for (int k1 = 0; k1 < 16; ++k1) {
                const float2 sv = l0 ? v[(32 - k1) & 31] : v[31 - k1];
                r[k1].x = __shfl_sync(0xffffffffu, sv.x, partner);
                r[k1].y = __shfl_sync(0xffffffffu, sv.y, partner);
              }
              cp_async_wait_all();                   // this frame's |S| row (issued between the two forward passes)
              __syncwarp();
            }
            const float* srow = reinterpret_cast<const float*>(buf) + s_off;
            float2 snd[16];
            float2 z512 = make_float2(0.0f, 0.0f);
#pragma unroll
            for (int k1 = 0; k1 < 16; ++k1) {
              const int k = 32 * k1 + lane, kp = 1024 - k;
              float2 Yk, Yp;
              if constexpr (MODE == MODE_GL_ITER) {
                const float Sk = spec_to_mag<SRC>(srow[k], g);
                const float Sp = spec_to_mag<SRC>(srow[kp], g);
                const float2 A = v[k1], B = r[k1];
                const float2 E2 = make_float2(A.x + B.x, A.y - B.y);
                const float2 D2 = make_float2(A.x - B.x, A.y + B.y);
                const float2 Tt = cmul(gt[k], D2);
                const float2 Xk = make_float2(E2.x + Tt.x, E2.y + Tt.y);          // 2 X[k]
                const float2 Xp = make_float2(E2.x - Tt.x, -(E2.y - Tt.y));       // 2 X[1024-k]
                const float mk = Xk.x * Xk.x + Xk.y * Xk.y;
                const float mp = Xp.x * Xp.x + Xp.y * Xp.y;
                const float ik = rsqrtf(fmaxf(mk, 1e-37f)), ip = rsqrtf(fmaxf(mp, 1e-37f));
                const float fk = Sk * ik, fp = Sp * ip;
                // Y = S X/|X|;  np.angle(0) = 0  ->  Y = S  (the imaginary part is 0 * finite = 0 already)
                Yk = make_float2(mk > 1e-37f ? Xk.x * fk : Sk, Xk.y * fk);
                Yp = make_float2(mp > 1e-37f ? Xp.x * fp : Sp, Xp.y * fp);
                if (SC && own) {
                  const float dk = 0.5f * mk * ik - Sk, dp = 0.5f * mp * ip - Sp;
                  sc_num += dk * dk + dp * dp;
                  sc_den += Sk * Sk + Sp * Sp;
                }
              } else if constexpr (SRC == SRC_COMPLEX) {
                // shifted coordinates: Y'[k] = Y[k] * exp(+j 2 pi k lpad / n_fft)
                const float2* in = reinterpret_cast<const float2*>(a.cplx_in) + row * kF;
                Yk = cmul(in[k], shift_phasor(k, lpad, 1.0f));
                Yp = cmul(in[kp], shift_phasor(kp, lpad, 1.0f));
              } else {
                // phases are given for the un-shifted frame: theta'[k] = theta[k] + 2 pi k lpad / n_fft
                const float Sk = spec_to_mag<SRC>(__ldg(a.spec + row * kF + k), g);
                const float Sp = spec_to_mag<SRC>(__ldg(a.spec + row * kF + kp), g);
                float sk, ck, sp, cp;
                if (a.angles != nullptr) {
                  sincosf(__ldg(a.angles + row * kF + k), &sk, &ck);
                  sincosf(__ldg(a.angles + row * kF + kp), &sp, &cp);
                } else {
                  sincospif(2.0f * philox_uniform(a.seed, (unsigned long long)(row * kF + k)), &sk, &ck);
                  sincospif(2.0f * philox_uniform(a.seed, (unsigned long long)(row * kF + kp)), &sp, &cp);
                }
                Yk = cmul(make_float2(Sk * ck, Sk * sk), shift_phasor(k, lpad, 1.0f));
                Yp = cmul(make_float2(Sp * cp, Sp * sp), shift_phasor(kp, lpad, 1.0f));
              }
              if (MODE == MODE_SYNTH && l0 && k1 == 0) { Yk.y = 0.0f; Yp.y = 0.0f; }   // irfft ignores Im of DC / Nyquist
              // Z'2[k] = P + Q, Z'2[1024-k] = conj(P - Q),  P = Y[k] + conj(Y[1024-k]),  Q = conj(G_k) (Y[k] - conj(Y[1024-k]))
              const float2 G = gt[k];
              const float2 P = make_float2(Yk.x + Yp.x, Yk.y - Yp.y);
              const float2 D = make_float2(Yk.x - Yp.x, Yk.y + Yp.y);
              const float2 Q = make_float2(G.x * D.x + G.y * D.y, G.x * D.y - G.y * D.x);
              v[k1] = make_float2(P.x + Q.x, -(P.y + Q.y));       // conj(Z'2[k])
              snd[k1] = make_float2(P.x - Q.x, P.y - Q.y);        // conj(Z'2[1024-k])
            }
            if (l0) {   // k = 512 (self-paired): X = conj(Z[512]), Z'2 = 2 conj(Y), conj(Z'2) = 2 Y
              float2 Y;
              if constexpr (MODE == MODE_GL_ITER) {
                const float S5 = spec_to_mag<SRC>(srow[512], g);
                const float2 X = make_float2(v[16].x, -v[16].y);
                const float m = X.x * X.x + X.y * X.y;
                const float im = rsqrtf(fmaxf(m, 1e-37f));
                const float f = S5 * im;
                Y = make_float2(m > 1e-37f ? X.x * f : S5, X.y * f);
                if (SC && own) {
                  const float d = m * im - S5;         // |X| = |Z[512]| (no factor 2 here)
                  sc_num += d * d;
                  sc_den += S5 * S5;
                }
              } else if constexpr (SRC == SRC_COMPLEX) {
                Y = cmul((reinterpret_cast<const float2*>(a.cplx_in) + row * kF)[512], shift_phasor(512, lpad, 1.0f));
              } else {
                const float S5 = spec_to_mag<SRC>(__ldg(a.spec + row * kF + 512), g);
                float s5, c5;
                if (a.angles != nullptr) sincosf(__ldg(a.angles + row * kF + 512), &s5, &c5);
                else sincospif(2.0f * philox_uniform(a.seed, (unsigned long long)(row * kF + 512)), &s5, &c5);
                Y = cmul(make_float2(S5 * c5, S5 * s5), shift_phasor(512, lpad, 1.0f));
              }
              z512 = make_float2(2.0f * Y.x, 2.0f * Y.y);
            }
            // hand the partner its half of each pair
#pragma unroll
            for (int k1 = 0; k1 < 16; ++k1) {
              r[k1].x = __shfl_sync(0xffffffffu, snd[k1].x, partner);
              r[k1].y = __shfl_sync(0xffffffffu, snd[k1].y, partner);
            }
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const float2 from_other = r[15 - i];                             // own register 31 - k1
              const float2 from_self = i == 0 ? z512 : r[16 - i];              // lane 0: own register 32 - k1
              v[16 + i] = l0 ? from_self : from_other;
            }
          }

          // ------------------------------------------------------------------ 1024-point transform, 32 x 32
          // v[n2] = z[lane + 32 n2]  ->  v[k1] = Z[32 k1 + lane]
#pragma unroll 1
          for (int pass = 0; pass < 2; ++pass) {
            fft32<false, 32>(v);
            if (pass == 0) {
#pragma unroll
              for (int k2 = 1; k2 < 32; ++k2) v[k2] = cmul(v[k2], tw[k2 * 32 + lane]);     // W_1024^(lane * k2)
              __syncwarp();
#pragma unroll
              for (int k2 = 0; k2 < 32; ++k2) buf[k2 * kRowStride + lane] = v[k2];
              __syncwarp();
#pragma unroll
              for (int n1 = 0; n1 < 32; n1 += 2) {
                const float4 q = *reinterpret_cast<const float4*>(&buf[lane * kRowStride + n1]);
                v[n1] = make_float2(q.x, q.y);
                v[n1 + 1] = make_float2(q.z, q.w);
              }
              if (MODE == MODE_GL_ITER && half == 0) {
                // the exchange buffer is idle until the inverse transform: stream this frame's |S| row into it now,
                // so that the HBM latency hides behind the second pass
                __syncwarp();
                s_off = row_to_smem_async(reinterpret_cast<float*>(buf), a.spec + row * kF, a.spec, a.spec_end, lane);
              }
            }
          }
        }  // halves

        if constexpr (MODE == MODE_ANALYSIS) {
          // ---------------------------------------------------------------- spectrum out
          // X[k] = (E2 + G_k D2)/2, X[1024-k] = conj(E2 - G_k D2)/2 ; undo the circular shift for complex output
          const int partner = (32 - lane) & 31;
          const bool l0 = lane == 0;
          const int lpad = (kNfft - g.win) >> 1;
          float* magbuf = reinterpret_cast<float*>(buf);
          float2* cout = reinterpret_cast<float2*>(a.cplx_out) + row * kF;
          __syncwarp();
#pragma unroll
          for (int k1 = 0; k1 < 16; ++k1) {
            const float2 sv = l0 ? v[(32 - k1) & 31] : v[31 - k1];
            float2 B;
            B.x = __shfl_sync(0xffffffffu, sv.x, partner);
            B.y = __shfl_sync(0xffffffffu, sv.y, partner);
            const int k = 32 * k1 + lane, kp = 1024 - k;
            const float2 A = v[k1];
            const float2 E2 = make_float2(A.x + B.x, A.y - B.y);
            const float2 D2 = make_float2(A.x - B.x, A.y + B.y);
            const float2 Tt = cmul(gt[k], D2);
            float2 Xk = make_float2(0.5f * (E2.x + Tt.x), 0.5f * (E2.y + Tt.y));
            float2 Xp = make_float2(0.5f * (E2.x - Tt.x), -0.5f * (E2.y - Tt.y));
            if (l0 && k1 == 0) { Xk.y = 0.0f; Xp.y = 0.0f; }
            if constexpr (SRC == OUT_COMPLEX) {
              // frame tap m sits at transform sample lpad + m: X_true[k] = X[k] * exp(-j 2 pi k lpad / n_fft)
              cout[k] = cmul(Xk, shift_phasor(k, lpad, -1.0f));
              cout[kp] = cmul(Xp, shift_phasor(kp, lpad, -1.0f));
            } else {
              magbuf[k] = sqrtf(Xk.x * Xk.x + Xk.y * Xk.y);
              magbuf[kp] = sqrtf(Xp.x * Xp.x + Xp.y * Xp.y);
            }
          }
          if (l0) {   // k = 512: X = conj(Z[512])
            const float2 Xc = make_float2(v[16].x, -v[16].y);
            if constexpr (SRC == OUT_COMPLEX) cout[512] = cmul(Xc, shift_phasor(512, lpad, -1.0f));
            else magbuf[512] = sqrtf(Xc.x * Xc.x + Xc.y * Xc.y);
          }
          if constexpr (SRC == OUT_FEATURES) {
            __syncwarp();
            if (a.lin_out != nullptr) {
              float* out = a.lin_out + row * kF;
              for (int k = lane; k < kF; k += 32) out[k] = amp_to_norm_db(magbuf[k], g);
            }
            if (a.mel_out != nullptr) {
              float* out = a.mel_out + row * g.num_mels;
              for (int m = lane; m < g.num_mels; m += 32) {
                const int lo = tb.mel_lo[m], cnt = tb.mel_cnt[m];
                const float* mv = tb.mel_val + m * tb.mel_ld;
                float acc = 0.0f;
                for (int c = 0; c < cnt; ++c) acc = fmaf(__ldg(mv + c), magbuf[lo + c], acc);
                out[m] = amp_to_norm_db(acc, g);
              }
            }
          }
        } else {
          // ---------------------------------------------------------------- window -> the warp's overlap-add slot
          // v[n2] = conj(z'[lane + 32 n2]):  y[2q] = Re, y[2q+1] = -Im
          __syncwarp();                       // all lanes done reading the exchange buffer: it becomes the slot
          float* slot = reinterpret_cast<float*>(buf);
#pragma unroll
          for (int n2 = 0; n2 < 32; ++n2) {
            const int q = lane + 32 * n2;
            if (n2 < NZ) {
              if (q < g.half)
                *reinterpret_cast<float2*>(slot + 2 * q) = make_float2(v[n2].x * wE[q], -v[n2].y * wO[q]);
              else if (kStd && 2 * q < ND * g.hop)
                *reinterpret_cast<float2*>(slot + 2 * q) = make_float2(0.0f, 0.0f);      // zero padding up to 5*hop
            } else if (kStd && 2 * q < ND * g.hop) {
              *reinterpret_cast<float2*>(slot + 2 * q) = make_float2(0.0f, 0.0f);
            }
          }
        }
      }

      __syncthreads();                             // slots complete; planes consumed
      const bool have_next = jt + 1 < jb;
      if constexpr (MODE != MODE_SYNTH) {
        if (have_next) stage_load(jt + 1);         // next span: loads in flight while this tile is overlap-added
      }

      if constexpr (MODE != MODE_ANALYSIS) {
        // -------------------------------------------------------------------- overlap-add + 1/(N wss) + store
        // each thread owns a residue rr (mod hop): acc[j] is span sample j*hop + rr; frame f adds its taps
        // rr + d*hop (d < ND) to acc[f + d].  All register indices are static; lanes read consecutive addresses.
        // The carry (samples that later frames still add to) is read and re-written by the same thread.
        const int fv_lo = first_needed > t0 ? first_needed - t0 : 0;
        const int fv_hi = (T - t0) < kNF ? (T - t0) : kNF;
        // the utterance's last tile also flushes what would be its carry (samples up to hop*(T-1) end there)
        const int out_len = (t0 + kNF >= T) ? g.span_len : kNF * g.hop;
        float* __restrict__ dst = a.wav_out + woff;
        const bool interior = kStd && write_out && fv_lo == 0 && fv_hi == kNF && i0 >= 0 && t0 + kNF < T &&
                              i0 + kNF * g.hop <= L && t0 >= ND - 1;
        for (int rr = tid; rr < g.hop; rr += kThreads) {
          float acc[kNF + ND - 1];
#pragma unroll
          for (int j = 0; j < kNF + ND - 1; ++j) {
            const int sidx = j * g.hop + rr;
            acc[j] = (has_carry && j < ND - 1 && sidx < g.carry_len) ? carry[sidx] : 0.0f;
          }
          if (interior) {
#pragma unroll
            for (int f = 0; f < kNF; ++f) {
              const float* sl = smem + f * kBufFloats + rr;
#pragma unroll
              for (int d = 0; d < ND; ++d) acc[f + d] += sl[d * g.hop];       // slots are zero padded to ND*hop
            }
            const float inv = pw[rr];
#pragma unroll
            for (int j = 0; j < kNF; ++j) dst[i0 + j * g.hop + rr] = acc[j] * inv;
#pragma unroll
            for (int j = kNF; j < kNF + ND - 1; ++j) {
              const int c = (j - kNF) * g.hop + rr;
              if (c < g.carry_len) carry[c] = acc[j];
            }
          } else {
#pragma unroll
            for (int f = 0; f < kNF; ++f) {
              if (f >= fv_lo && f < fv_hi) {
                const float* sl = smem + f * kBufFloats + rr;
#pragma unroll
                for (int d = 0; d < ND; ++d)
                  if (kStd || rr + d * g.hop < g.win) acc[f + d] += sl[d * g.hop];
              }
            }
            const int dmax = (g.win - 1 - rr) / g.hop;
#pragma unroll 1
            for (int j = 0; j < kNF + ND - 1; ++j) {
              const int sidx = j * g.hop + rr;
              float val = 0.0f;
#pragma unroll
              for (int jj = 0; jj < kNF + ND - 1; ++jj) val = (jj == j) ? acc[jj] : val;
              if (sidx >= g.span_len) continue;
              if (sidx < out_len) {
                const int i = i0 + sidx;
                if (write_out && i >= 0 && i < L) {
                  float inv = pw[rr];
                  if (t0 + j - dmax < 0 || t0 + j > T - 1) {          // some overlapping frame does not exist
                    float ws = 0.0f;
                    for (int d = 0, m = rr; m < g.win; ++d, m += g.hop) {
                      const int tt = t0 + j - d;
                      if (tt >= 0 && tt < T) {
                        const float wv = (m & 1) ? wO[m >> 1] : wE[m >> 1];
                        ws = fmaf(wv, wv, ws);
                      }
                    }
                    inv = ws > 1.17549435e-38f ? kInvN / ws : kInvN;   // librosa: divide only where wss > tiny
                  }
                  dst[i] = val * inv;
                }
              } else {
                carry[sidx - out_len] = val;
              }
            }
          }
        }
        has_carry = true;
      }

      if constexpr (MODE != MODE_SYNTH) {
        if (have_next) stage_store(jt + 1);
      }
      __syncthreads();                             // planes of the next tile ready; slots and carry settled
    }  // tiles of the segment

    if constexpr (MODE == MODE_GL_ITER && SC) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        sc_num += __shfl_xor_sync(0xffffffffu, sc_num, o);
        sc_den += __shfl_xor_sync(0xffffffffu, sc_den, o);
      }
      if (lane == 0) {
        atomicAdd(a.sc_acc + 2 * u, sc_num);
        atomicAdd(a.sc_acc + 2 * u + 1, sc_den);
      }
    }
  }  // segments
}

}  // namespace ttsa
