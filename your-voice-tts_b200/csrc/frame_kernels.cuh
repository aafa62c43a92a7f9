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
// conj(Z') and the synthesis window (applied inside the overlap-add) carries the sign of the imaginary parts.  The
// transform body therefore exists once in the instruction stream (a rolled loop of four 32-point passes), which keeps
// the kernel inside the instruction cache.
//
// All per-frame arithmetic runs on packed fp32 pairs (FFMA2/FADD2/FMUL2, fft32p.cuh): a thread's 32 complex values are
// 16 (re, re) + 16 (im, im) register pairs.  The FP32 lane work is unchanged, but the packed form halves the issue
// slots it needs, and the kernel is issue-bound (measured: FP pipe 39 % busy at 64 % issue utilisation before packing).
//
// Template parameters HOP / WIN fix the geometry at compile time for the shipped configurations (every shared-memory
// offset and bound becomes an immediate; Layout in common.cuh is shared with the host); MOM adds the previous estimate
// for the opt-in fast Griffin-Lim.  The kernel sits at its 128-register limit (2 CTAs x 8 warps per SM): DESIGN.md
// section 4 lists the reorderings and extra live values that were measured and cost 1-4 % each.
#pragma once
#include "common.cuh"
#include "fft32p.cuh"

// Phase-skipping timing probes exist only in profiling builds (build.py with TTSA_NVCC_EXTRA=-DTTSA_PROFILE_BUILD): the
// shipped library cannot be told to skip work.
#ifdef TTSA_PROFILE_BUILD
#define TTSA_SKIP(a, bit) (((a).debug & (bit)) != 0)
#else
#define TTSA_SKIP(a, bit) false
#endif

namespace ttsa {

enum { MODE_GL_ITER = 0, MODE_SYNTH = 1, MODE_ANALYSIS = 2 };
enum { SRC_MAG = 0, SRC_NORM_DB = 1, SRC_COMPLEX = 2 };   // GL_ITER / SYNTH input kind
enum { OUT_COMPLEX = 0, OUT_FEATURES = 1 };               // ANALYSIS output kind

struct FrameArgs {
  const float* spec;      // [sum_T, F]      GL_ITER / SYNTH (mag or normalised dB)
  const float* spec_end;  // one past the last element of spec (bounds for the 16-byte row copies)
  long long rows_total;   // rows of the spectrogram-domain tensors (bounds for angles / complex row copies)
  const float* angles;    // [sum_T, F]      SYNTH: initial phases in radians, or nullptr -> counter RNG
  const float* cplx_in;   // [sum_T, F, 2]   SYNTH with SRC_COMPLEX
  const float* wav_in;    // packed wav      GL_ITER (previous y) / ANALYSIS
  const float* wav_end;   // one past the last float of wav_in (bounds for 16-byte span copies)
  float* wav_out;         // packed wav      GL_ITER / SYNTH
  float* cplx_out;        // [sum_T, F, 2]   ANALYSIS OUT_COMPLEX
  float* lin_out;         // [sum_T, F]      ANALYSIS OUT_FEATURES (nullable)
  float* mel_out;         // [sum_T, mels]   ANALYSIS OUT_FEATURES (nullable)
  float* sc_acc;          // [B, 2]          GL_ITER with SC: (sum (|X|-S)^2, sum S^2)
  unsigned long long seed;
  int preemph;            // ANALYSIS: apply y[n] - p*y[n-1] while staging
  const float* wav_prev;  // GL_ITER with momentum: the estimate before wav_in; the kernel transforms wav_in - beta * wav_prev
  float beta;             //   beta = momentum / (1 + momentum)   (fast Griffin-Lim, opt-in; not in the reference)
  int ola_phase, ola_phases;  // any-size path (generic_kernels.cuh): this launch overlap-adds the frames t = ola_phase (mod
                          //   ola_phases); frames of one phase do not overlap, so no atomics and a fixed summation order
  int* wps_flags;         // warp-stream GL_ITER (gl_stream.cuh): per-warp "head zone stored" flag, compared with wps_epoch
  int wps_epoch;          //   (iteration number; ttsa_griffin_lim zeroes the flags once per call)
  int* wps_done;          //   per-warp "iteration complete" flag (same epochs) for launches that run several iterations
  int wps_iters;          //   iterations this launch runs: iteration j reads wav_in / wav_out for even / odd j, writes the other
#ifdef TTSA_WPS_TRACE
  unsigned long long* wps_trace;   // [warps][64 iterations][8 stamps] (experiment builds)
#endif
  int debug;              // profiling builds only (-DTTSA_PROFILE_BUILD + env TTSA_DEBUG): 1 = skip the frame phase,
                          // 2 = skip overlap-add + staging work; the shipped library compiles these branches out
};

// ---------------------------------------------------------------------------------------------------------
// counter-based RNG (Philox4x32-10) for the initial phases when none are injected
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float4 philox_uniform4(unsigned long long seed, unsigned long long ctr) {
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
  constexpr float s = 1.0f / 16777216.0f;   // [0, 1)
  return make_float4((float)(c0 >> 8) * s, (float)(c1 >> 8) * s, (float)(c2 >> 8) * s, (float)(c3 >> 8) * s);
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

// Asynchronous copy of `n` consecutive floats starting at `ptr` (4-byte aligned) into dst with 16-byte chunks taken
// from the enclosing 16-byte-aligned range: element k lands at dst[off + k], off = returned value (0..3).  Chunks that
// would touch memory outside [lo, hi) fall back to 4-byte copies of the in-range elements.
__device__ __forceinline__ int span_to_smem_async(float* dst, const float* ptr, int n, const float* lo, const float* hi, int lane) {
  const int off = (int)((reinterpret_cast<unsigned long long>(ptr) >> 2) & 3ull);
  const float* base = ptr - off;                      // 16-byte aligned
  const int nchunks = (off + n + 3) >> 2;
  if (base >= lo && base + 4 * nchunks <= hi) {       // every chunk inside the tensor (all rows but the first / last)
    for (int c = lane; c < nchunks; c += 32) cp_async16(dst + 4 * c, base + 4 * c);
    return off;
  }
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
// The same for a compile-time length: the chunk loop is fully unrolled (one predicated LDGSTS per 32 chunks).
template <int N>
__device__ __forceinline__ int span_to_smem_async_n(float* dst, const float* ptr, const float* lo, const float* hi, int lane) {
  const int off = (int)((reinterpret_cast<unsigned long long>(ptr) >> 2) & 3ull);
  const float* base = ptr - off;
  const int nchunks = (off + N + 3) >> 2;
  constexpr int kMaxChunks = (3 + N + 3) >> 2;
  if (base >= lo && base + 4 * kMaxChunks <= hi) {
    const float* g = base + 4 * lane;
    float* d = dst + 4 * lane;
#pragma unroll
    for (int k = 0; k < (kMaxChunks + 31) / 32; ++k)
      if (32 * k + 31 < (N >> 2) || lane + 32 * k < nchunks) cp_async16(d + 128 * k, g + 128 * k);
    return off;
  }
  return span_to_smem_async(dst, ptr, N, lo, hi, lane);
}
// one spectrogram row (kF floats)
__device__ __forceinline__ int row_to_smem_async(float* dst, const float* row_ptr, const float* lo, const float* hi, int lane) {
  return span_to_smem_async(dst, row_ptr, kF, lo, hi, lane);
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
  // lg2.approx: absolute error ~1e-7 in log2, i.e. ~1e-6 dB -- two orders below the float32 conditioning of |X| itself
  // (.ftz: the argument is at least min_amp = 10^(min_level_db / 20), a normal number for any sensible config, so the three
  // instructions with which __log2f rescales denormal arguments are dead weight -- 36 calls per frame in the feature kernels)
  float lg;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lg) : "f"(fmaxf(g.min_amp, a)));
  const float v = fmaf(g.n_a, lg, g.n_b);
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
__device__ __forceinline__ float rsqrt_fast(float x) {
  float y;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float sqrt_fast(float x) {
  float y;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float2 shfl2(float a, float b, int srclane) {
  return make_float2(__shfl_sync(0xffffffffu, a, srclane), __shfl_sync(0xffffffffu, b, srclane));
}

// HOP, WIN > 0: geometry fixed at compile time (the shipped configurations); 0: read from Geo at run time.
// FINE (Griffin-Lim iteration on batches too small to fill the GPU, e.g. the server's one sentence at a time,
// server/synthesizer.py:147-157): one CTA per "fine segment" -- the utterance's first 8 frames, then 8 - nwarm owned frames
// each -- and ONE tile per CTA whose first nwarm frames are the recomputed neighbours.  A 6 s utterance then occupies 120
// CTAs for one tile phase per iteration instead of 61 CTAs for two (warm-up tile + own tile).
template <int MODE, int SRC, int NZ, bool SC, int HOP = 0, int WIN = 0, bool MOM = false, bool FINE = false>
__global__ void __launch_bounds__(kThreads, 2)
frame_kernel(const Geo g, const Tables tb, const BatchDev bd, const FrameArgs a) {
  extern __shared__ __align__(16) float smem[];
  constexpr bool kFixed = HOP > 0;
  constexpr Layout kLy = make_layout(kFixed ? HOP : 256, kFixed ? WIN : 1024, NZ);
  const Layout ly = kFixed ? kLy : g.ly;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr float kInvN = 1.0f / (float)kNfft;
  constexpr float kTiny = 1e-37f;
  constexpr float kPhaseEps = 1e-18f;
  // kernel class: NZ 20 = "standard" geometry (<= 20 non-zero packed rows, <= 5 window taps per residue mod hop);
  // NZ 32 = anything up to win <= 9*hop.
  constexpr bool kStd = (NZ <= 20);
  constexpr int ND = kStd ? 5 : kNF + 1;
  constexpr int kStage = 13;                      // span samples per thread prefetched through registers (13*256 >= 7*hop+win
                                                  // for every shipped geometry; any remainder is copied synchronously)

  float* const buf = smem + warp * kBufFloats;
  float* const plane0 = smem + ly.sm_plane0;
  float* const plane1 = smem + ly.sm_plane1;
  float* const carry = smem + ly.sm_carry0;
  const float2* const wE2 = reinterpret_cast<const float2*>(smem + ly.sm_wE);
  const float2* const wO2 = reinterpret_cast<const float2*>(smem + ly.sm_wO);
  float* const pw = smem + ly.sm_pw;
  float* const wsyn = smem + ly.sm_wsyn;         // synthesis window per tap, odd taps negated (conjugate-FFT inverse)
  const float4* const tw4 = reinterpret_cast<const float4*>(smem + ly.sm_tw);
  const float4* const g4 = reinterpret_cast<const float4*>(smem + ly.sm_g);

  // plan tables -> shared memory, once per persistent CTA: ONE bulk asynchronous copy of the host-built image
  // (paired windows, 1/(N wss), signed synthesis window, twiddles) signalled through an mbarrier, while the threads zero
  // the staged-span planes (their tails stay zero for the kernel's lifetime)
  const unsigned mbar = (unsigned)__cvta_generic_to_shared(smem + ly.sm_mbar);
  if (tid == 0) {
    const unsigned dst = (unsigned)__cvta_generic_to_shared(smem + ly.sm_wE);
    const unsigned bytes = (unsigned)ly.image_floats * 4u;
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(tb.smem_image), "r"(bytes), "r"(mbar) : "memory");
  }
  for (int i = tid; i < ly.plane_len; i += kThreads) { plane0[i] = 0.0f; plane1[i] = 0.0f; }
  if constexpr (MODE == MODE_ANALYSIS && SRC == OUT_FEATURES) {
    for (int i = tid; i < g.mel_smem_floats; i += kThreads) smem[ly.sm_total + i] = tb.mel_compact[i];
  }
  __syncthreads();                                 // the mbarrier is initialised for everyone
  {
    unsigned done = 0;
    while (!done) {
      asm volatile("{\n\t.reg .pred q;\n\tmbarrier.try_wait.parity.shared::cta.b64 q, [%1], 0;\n\tselp.u32 %0, 1, 0, q;\n\t}\n"
                   : "=r"(done) : "r"(mbar) : "memory");
    }
  }

  // contiguous tile range of this CTA over the flattened (utterance, tile) list.  Everything that depends only on
  // the batch layout and the (constant) spectrogram happens BEFORE the dependency wait below: the utterance search is a
  // chain of dependent global loads, and the first tile's |S| rows can already travel towards L2.
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const long long tile_lo = FINE ? (long long)blockIdx.x : (long long)blockIdx.x * bd.total_tiles / gridDim.x;
  const long long tile_hi = FINE ? (long long)blockIdx.x + 1 : (long long)(blockIdx.x + 1) * bd.total_tiles / gridDim.x;
  const int* const seg_off = FINE ? bd.fine_off : bd.tile_off;      // prefix sums of the work list this launch walks
  int u = 0;
  {
    int lo = 0, hi = bd.B;   // largest u with seg_off[u] <= tile_lo
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (seg_off[mid] <= tile_lo) lo = mid; else hi = mid;
    }
    u = lo;
  }
  if constexpr (MODE == MODE_GL_ITER && !FINE) {
    if (tile_lo < tile_hi) {
      const int t_first = (int)(tile_lo - bd.tile_off[u]) * kNF;
      const int n_rows = min(kNF, bd.T[u] - t_first);
      const char* rows = reinterpret_cast<const char*>(a.spec + (bd.frame_off[u] + t_first) * kF);
      const int bytes = n_rows * kF * 4;
      for (int o = tid * 128; o < bytes; o += kThreads * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(rows + o));
    }
  }
  // programmatic dependent launch: wait here until the previous kernel's global writes (the waveform this kernel
  // reads, the buffer it overwrites) are complete
  asm volatile("griddepcontrol.wait;" ::: "memory");

  long long tile = tile_lo;
  while (tile < tile_hi) {
    while (u + 1 < bd.B && tile >= seg_off[u + 1]) ++u;
    const int T = bd.T[u];
    const int toff = seg_off[u];
    const long long seg_end = tile_hi < (long long)seg_off[u + 1] ? tile_hi : (long long)seg_off[u + 1];
    // fine segment s of the utterance: owned frames [sfa, sfb), tile origin t0f (its first sfa - t0f frames are warm-up)
    const int fs = (int)(tile - toff);
    const int fown = kNF - ly.nwarm;
    const int sfa = FINE ? (fs == 0 ? 0 : kNF + (fs - 1) * fown) : 0;
    const int sfb = FINE ? min(T, fs == 0 ? kNF : sfa + fown) : T;
    const int t0f = FINE ? (fs == 0 ? 0 : sfa - ly.nwarm) : 0;
    const int ja = FINE ? 0 : (int)(tile - toff);
    const int jb = FINE ? 1 : (int)(seg_end - toff);
    tile = seg_end;
    const int L = bd.wav_len[u];                 // signal length (GL / SYNTH: hop*(T-1))
    if (MODE != MODE_ANALYSIS && L <= 0) continue;
    const long long frow0 = bd.frame_off[u];
    const long long woff = bd.wav_off[u];
    const float* __restrict__ src = (MODE != MODE_SYNTH) ? a.wav_in + woff : nullptr;
    const float* __restrict__ srcp = MOM ? a.wav_prev + woff : nullptr;     // fast Griffin-Lim: previous estimate

    // frames of earlier tiles still overlap this segment's first owned sample: recompute them (no output)
    const bool warm = !FINE && (MODE != MODE_ANALYSIS) && ja > 0 && ly.nwarm > 0;
    const int first_needed = FINE ? t0f : (warm ? ja * kNF - ly.nwarm : 0);
    const int jt_first = warm ? ja - 1 : ja;
    bool has_carry = false;
    float sc_num = 0.0f, sc_den = 0.0f;

    // ---- span staging helpers: global -> registers -> (later) shared, de-interleaved by parity so that the
    //      re/im frame loads are conflict-free
    float stg[kStage];
    auto stage_load = [&](int jt) {
      const int i0 = (FINE ? t0f : jt * kNF) * ly.hop - ly.off0;
      if (i0 >= 1 && i0 + ly.span_len <= L) {                       // interior span: no reflection
        const float* __restrict__ sp = src + i0;
#pragma unroll
        for (int e = 0; e < kStage; ++e) {
          const int s = tid + e * kThreads;
          float val = 0.0f;
          if (s < ly.span_len) {
            val = __ldg(sp + s);
            if constexpr (MOM) val = fmaf(-a.beta, __ldg(srcp + i0 + s), val);
            if (MODE == MODE_ANALYSIS && a.preemph) val = fmaf(-g.preemph, __ldg(sp + s - 1), val);
          }
          stg[e] = val;
        }
      } else {
#pragma unroll
        for (int e = 0; e < kStage; ++e) {
          const int s = tid + e * kThreads;
          float val = 0.0f;
          if (s < ly.span_len) {
            const int j = reflect_index(i0 + s, L);
            val = __ldg(src + j);
            if constexpr (MOM) val = fmaf(-a.beta, __ldg(srcp + j), val);
            if (MODE == MODE_ANALYSIS && a.preemph) val = fmaf(-g.preemph, j > 0 ? __ldg(src + j - 1) : 0.0f, val);
          }
          stg[e] = val;
        }
      }
    };
    auto stage_store = [&](int jt) {
      float* const pl = ((tid & 1) ? plane1 : plane0) + (tid >> 1);    // parity of tid + e*kThreads = parity of tid
#pragma unroll
      for (int e = 0; e < kStage; ++e)
        if (tid + e * kThreads < ly.span_len) pl[e * (kThreads / 2)] = stg[e];
      const int i0 = (FINE ? t0f : jt * kNF) * ly.hop - ly.off0;
      for (int s = tid + kStage * kThreads; s < ly.span_len; s += kThreads) {     // spans longer than the register window
        const int j = reflect_index(i0 + s, L);
        float val = __ldg(src + j);
        if constexpr (MOM) val = fmaf(-a.beta, __ldg(srcp + j), val);
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
      const int t0 = FINE ? t0f : jt * kNF;
      const int i0 = t0 * ly.hop - ly.off0;        // sample index of span position 0
      const bool write_out = jt >= ja;

      if constexpr (MODE != MODE_SYNTH) {
        // pull the next tile's span towards L2 while this tile is being transformed
        if (jt + 1 < jb) {
          const int pi = (jt + 1) * kNF * ly.hop - ly.off0 + tid * 32;
          if (pi >= 0 && pi < L && tid * 32 < ly.span_len) asm volatile("prefetch.global.L2 [%0];" ::"l"(src + pi));
        }
      }
      const int t = t0 + warp;
      if (t < sfb && t >= first_needed && !TTSA_SKIP(a, 1)) {
        const long long row = frow0 + t;
        const bool own = FINE ? t >= sfa : t >= ja * kNF;   // warm-up frames are copies of another segment's frames
        // 32 complex values per thread as packed pairs: R[m] = (re[2m], re[2m+1]), I[m] = (im[2m], im[2m+1])
        float2 R[16], I[16];
        int s_off = 0;
        constexpr int kHalfBegin = (MODE == MODE_SYNTH) ? 1 : 0;
        constexpr int kHalfEnd = (MODE == MODE_ANALYSIS) ? 1 : 2;
        const int partner = (32 - lane) & 31;
        const bool l0 = lane == 0;
        const int lpad = (kNfft - ly.win) >> 1;

#pragma unroll 1
        for (int half = kHalfBegin; half < kHalfEnd; ++half) {
          if (half == 0) {
            // ---------------------------------------------------------------- window the frame
            // element n2 of this lane is z[lane + 32 n2] = x[2q] + j x[2q+1], q = lane + 32 n2; rows past the
            // window read zero window taps (the planes' tails are zero, so the products are exact zeros)
            if constexpr (MODE != MODE_SYNTH) {
              const int o = warp * ly.hop;
              const float* re_p = ((o & 1) ? plane1 + ((o - 1) >> 1) : plane0 + (o >> 1)) + lane;
              const float* im_p = ((o & 1) ? plane0 + ((o + 1) >> 1) : plane1 + (o >> 1)) + lane;
#pragma unroll
              for (int m = 0; m < 16; ++m) {
                if (2 * m < NZ) {
                  R[m] = __fmul2_rn(make_float2(re_p[64 * m], re_p[64 * m + 32]), wE2[m * 32 + lane]);
                  I[m] = __fmul2_rn(make_float2(im_p[64 * m], im_p[64 * m + 32]), wO2[m * 32 + lane]);
                } else {
                  R[m] = make_float2(0.0f, 0.0f);
                  I[m] = make_float2(0.0f, 0.0f);
                }
              }
            }
          } else if constexpr (MODE != MODE_ANALYSIS) {
            // ---------------------------------------------------------------- per-bin step -> conj(Z')
            // lane holds Z[32 k1 + lane]; bin k pairs with 1024 - k, held by lane (32 - lane) & 31 in register
            // 31 - k1 (lane 0: its own register 32 - k1).  Each lane processes the 16 pairs whose first member is its
            // own register k1 < 16, two pairs (k1 = 2m, 2m+1) per packed instruction.
            float2 BR[8], BI[8];
            if constexpr (MODE == MODE_GL_ITER) {
              static_for<0, 8>([&](auto mc) {
                constexpr int m = decltype(mc)::value;
                constexpr int ms = (m == 0) ? 0 : 16 - m;       // lane 0, k1 = 2m: own register (32 - 2m) & 31 = 2 * ms
                const float s0r = l0 ? R[ms].x : R[15 - m].y, s0i = l0 ? I[ms].x : I[15 - m].y;
                const float s1r = l0 ? R[15 - m].y : R[15 - m].x, s1i = l0 ? I[15 - m].y : I[15 - m].x;
                BR[m] = shfl2(s0r, s1r, partner);
                BI[m] = shfl2(s0i, s1i, partner);
              });
              cp_async_wait_all();                   // this frame's |S| row (issued between the two forward passes)
              __syncwarp();
            }
            int a_off = 0;
            if constexpr (MODE == MODE_SYNTH) {
              // stage this frame's input rows in the (idle) exchange buffer: one exposed HBM round trip per frame
              // instead of one per group of bins
              __syncwarp();
              if constexpr (SRC == SRC_COMPLEX) {
                s_off = span_to_smem_async(buf, a.cplx_in + row * (2 * kF), 2 * kF, a.cplx_in,
                                           a.cplx_in + a.rows_total * (2 * kF), lane);
              } else {
                s_off = row_to_smem_async(buf, a.spec + row * kF, a.spec, a.spec_end, lane);
                if (a.angles != nullptr)
                  a_off = 1040 + row_to_smem_async(buf + 1040, a.angles + row * kF, a.angles, a.angles + a.rows_total * kF, lane);
              }
              cp_async_wait_all();
              __syncwarp();
            }
            const float* srow = buf + s_off;
            const float* arow = buf + a_off;
            float2 SR[8], SI[8];
            float2 z512 = make_float2(0.0f, 0.0f);
            static_for<0, 8>([&](auto mc) {
              constexpr int m = decltype(mc)::value;
              const int k0 = 64 * m + lane;                      // bins k0 (k1 = 2m) and k0 + 32 (k1 = 2m+1); partners 1024 - k
              const float4 gq = g4[m * 32 + lane];
              const float2 GX = make_float2(gq.x, gq.y), GY = make_float2(gq.z, gq.w);
              float2 YkR, YkI, YpR, YpI;
              if constexpr (MODE == MODE_GL_ITER) {
                const float2 Sk = make_float2(spec_to_mag<SRC>(srow[k0], g), spec_to_mag<SRC>(srow[k0 + 32], g));
                const float2 Sp = make_float2(spec_to_mag<SRC>(srow[1024 - k0], g), spec_to_mag<SRC>(srow[992 - k0], g));
                const float2 E2R = __fadd2_rn(R[m], BR[m]), E2I = __fadd2_rn(I[m], neg2(BI[m]));
                const float2 D2R = __fadd2_rn(R[m], neg2(BR[m])), D2I = __fadd2_rn(I[m], BI[m]);
                // 2 X[k] = E2 + G D2 ;  2 X[1024-k] = conj(E2 - G D2) = conj(2 E2 - 2 X[k])
                float2 XkR = __ffma2_rn(GX, D2R, E2R);
                XkR = __ffma2_rn(neg2(GY), D2I, XkR);
                float2 XkI = __ffma2_rn(GX, D2I, E2I);
                XkI = __ffma2_rn(GY, D2R, XkI);
                float2 XpR = __ffma2_rn(E2R, splat(2.0f), neg2(XkR));
                const float2 XpI = __ffma2_rn(E2I, splat(-2.0f), XkI);
                // Y = S X/|X| with np.angle(0) = 0, i.e. Y = S where X == 0: a real offset far below the rounding
                // error of any computed bin (2X is O(1e-7 * frame peak) at best) gives exact zeros the phase 0 without
                // a select, and the floor inside |2X|^2 keeps the reciprocal square root finite.  (Only a frame whose
                // WHOLE spectrum lies below ~1e-13, i.e. -260 dB, would see its phases biased by the offset.)
                XkR = __fadd2_rn(XkR, splat(kPhaseEps));
                XpR = __fadd2_rn(XpR, splat(kPhaseEps));
                const float2 mk = __ffma2_rn(XkI, XkI, __ffma2_rn(XkR, XkR, splat(kTiny)));
                const float2 mp = __ffma2_rn(XpI, XpI, __ffma2_rn(XpR, XpR, splat(kTiny)));
                const float2 ik = make_float2(rsqrt_fast(mk.x), rsqrt_fast(mk.y));
                const float2 ip = make_float2(rsqrt_fast(mp.x), rsqrt_fast(mp.y));
                const float2 fk = __fmul2_rn(Sk, ik), fp = __fmul2_rn(Sp, ip);
                YkR = __fmul2_rn(XkR, fk); YkI = __fmul2_rn(XkI, fk);
                YpR = __fmul2_rn(XpR, fp); YpI = __fmul2_rn(XpI, fp);
                if (SC && own) {
                  const float2 dk = __ffma2_rn(__fmul2_rn(mk, ik), splat(0.5f), neg2(Sk));   // |X| - S
                  const float2 dp = __ffma2_rn(__fmul2_rn(mp, ip), splat(0.5f), neg2(Sp));
                  sc_num += dk.x * dk.x + dk.y * dk.y + dp.x * dp.x + dp.y * dp.y;
                  sc_den += Sk.x * Sk.x + Sk.y * Sk.y + Sp.x * Sp.x + Sp.y * Sp.y;
                }
              } else {
                // SYNTH: Y given directly (complex input, or magnitude with injected / generated phases), expressed in
                // the shifted frame:  Y'[k] = Y[k] exp(+j 2 pi k lpad / n_fft)
                // Generated phases (no injected angles): one Philox4x32-10 call per (frame, m, lane) yields the four
                // uniforms of this step's four bins.  A uniformly random phase stays uniformly random under the
                // deterministic shift ramp, so generated phases are defined directly in the shifted frame.
                float2 y[4];
                float u4[4] = {0.0f, 0.0f, 0.0f, 0.0f};
                const bool gen = SRC != SRC_COMPLEX && a.angles == nullptr;
                if (gen) {
                  const float4 q = philox_uniform4(a.seed, ((unsigned long long)row * 16ull + (unsigned long long)m) * 32ull + (unsigned long long)lane);
                  u4[0] = q.x; u4[1] = q.y; u4[2] = q.z; u4[3] = q.w;
                }
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const int k = (e & 1) ? 1024 - (k0 + 32 * (e >> 1)) : k0 + 32 * (e >> 1);   // k0, 1024-k0, k0+32, 992-k0
                  if constexpr (SRC == SRC_COMPLEX) {
                    y[e] = cmul(make_float2(srow[2 * k], srow[2 * k + 1]), shift_phasor(k, lpad, 1.0f));
                  } else {
                    const float S = spec_to_mag<SRC>(srow[k], g);
                    float sn, cs;
                    if (gen) {
                      __sincosf(6.2831853071795864769f * u4[e], &sn, &cs);
                      y[e] = make_float2(S * cs, S * sn);
                    } else {
                      sincosf(arow[k], &sn, &cs);
                      y[e] = cmul(make_float2(S * cs, S * sn), shift_phasor(k, lpad, 1.0f));
                    }
                  }
                }
                if (m == 0 && l0) { y[0].y = 0.0f; y[1].y = 0.0f; }     // irfft ignores Im of DC / Nyquist
                YkR = make_float2(y[0].x, y[2].x); YkI = make_float2(y[0].y, y[2].y);
                YpR = make_float2(y[1].x, y[3].x); YpI = make_float2(y[1].y, y[3].y);
              }
              // Z'2[k] = P + Q, Z'2[1024-k] = conj(P - Q),  P = Y[k] + conj(Y[1024-k]),  Q = conj(G_k) (Y[k] - conj(Y[1024-k]))
              const float2 PR = __fadd2_rn(YkR, YpR), PI = __fadd2_rn(YkI, neg2(YpI));
              const float2 DR = __fadd2_rn(YkR, neg2(YpR)), DI = __fadd2_rn(YkI, YpI);
              float2 vR = __ffma2_rn(GX, DR, PR);                  // Re(P + Q),  Re Q = GX DR + GY DI
              vR = __ffma2_rn(GY, DI, vR);
              float2 vI = __ffma2_rn(neg2(GX), DI, neg2(PI));      // -Im(P + Q), Im Q = GX DI - GY DR
              vI = __ffma2_rn(GY, DR, vI);
              R[m] = vR; I[m] = vI;                                // conj(Z'2[k])
              SR[m] = __ffma2_rn(PR, splat(2.0f), neg2(vR));       // conj(Z'2[1024-k]) = P - Q
              SI[m] = __ffma2_rn(PI, splat(2.0f), vI);
            });
            if (l0) {   // k = 512 (self-paired, register 16 = R[8].x): X = conj(Z[512]), Z'2 = 2 conj(Y), conj(Z'2) = 2 Y
              float2 Y;
              if constexpr (MODE == MODE_GL_ITER) {
                const float S5 = spec_to_mag<SRC>(srow[512], g);
                const float2 X = make_float2(R[8].x, -I[8].x);
                const float m = X.x * X.x + X.y * X.y;
                const float im = rsqrt_fast(fmaxf(m, kTiny));
                const float f = S5 * im;
                Y = make_float2(m > kTiny ? X.x * f : S5, X.y * f);
                if (SC && own) {
                  const float d = m * im - S5;         // |X| = |Z[512]| (no factor 2 here)
                  sc_num += d * d;
                  sc_den += S5 * S5;
                }
              } else if constexpr (SRC == SRC_COMPLEX) {
                Y = cmul(make_float2(srow[1024], srow[1025]), shift_phasor(512, lpad, 1.0f));
              } else {
                const float S5 = spec_to_mag<SRC>(srow[512], g);
                float s5, c5;
                if (a.angles != nullptr) {
                  sincosf(arow[512], &s5, &c5);
                  Y = cmul(make_float2(S5 * c5, S5 * s5), shift_phasor(512, lpad, 1.0f));
                } else {
                  __sincosf(6.2831853071795864769f * philox_uniform4(a.seed, ((unsigned long long)row * 16ull + 8ull) * 32ull).x, &s5, &c5);
                  Y = make_float2(S5 * c5, S5 * s5);
                }
              }
              z512 = make_float2(2.0f * Y.x, 2.0f * Y.y);
            }
            // hand the partner its half of each pair: received r[k1] goes to own register 31 - k1 (lane 0: 32 - k1)
            float rr_[16], ri_[16];
            static_for<0, 8>([&](auto mc) {
              constexpr int m = decltype(mc)::value;
              rr_[2 * m] = __shfl_sync(0xffffffffu, SR[m].x, partner);
              rr_[2 * m + 1] = __shfl_sync(0xffffffffu, SR[m].y, partner);
              ri_[2 * m] = __shfl_sync(0xffffffffu, SI[m].x, partner);
              ri_[2 * m + 1] = __shfl_sync(0xffffffffu, SI[m].y, partner);
            });
            static_for<0, 8>([&](auto jc) {
              constexpr int j = decltype(jc)::value;       // registers 16 + 2j, 17 + 2j
              const float ar = (j == 0) ? z512.x : rr_[(16 - 2 * j) & 15], ai = (j == 0) ? z512.y : ri_[(16 - 2 * j) & 15];
              R[8 + j] = make_float2(l0 ? ar : rr_[15 - 2 * j], l0 ? rr_[15 - 2 * j] : rr_[14 - 2 * j]);
              I[8 + j] = make_float2(l0 ? ai : ri_[15 - 2 * j], l0 ? ri_[15 - 2 * j] : ri_[14 - 2 * j]);
            });
          }

          // ------------------------------------------------------------------ 1024-point transform, 32 x 32
          // element n2 = z[lane + 32 n2]  ->  element k1 = Z[32 k1 + lane]
#pragma unroll 1
          for (int pass = 0; pass < 2; ++pass) {
            fft32p(R, I);
            if (pass == 0) {
#pragma unroll
              for (int m = 0; m < 16; ++m) {                      // times W_1024^(lane * k2), k2 = 2m, 2m+1
                const float4 w = tw4[m * 32 + lane];
                const float2 WR = make_float2(w.x, w.y), WI = make_float2(w.z, w.w);
                const float2 nr = __ffma2_rn(R[m], WR, neg2(__fmul2_rn(I[m], WI)));
                I[m] = __ffma2_rn(R[m], WI, __fmul2_rn(I[m], WR));
                R[m] = nr;
              }
              __syncwarp();
#pragma unroll
              for (int m = 0; m < 16; ++m) {                      // row k2: [re 0..31 | im 0..31], column = lane
                buf[(2 * m) * kRowFloats + lane] = R[m].x;
                buf[(2 * m) * kRowFloats + 32 + lane] = I[m].x;
                buf[(2 * m + 1) * kRowFloats + lane] = R[m].y;
                buf[(2 * m + 1) * kRowFloats + 32 + lane] = I[m].y;
              }
              __syncwarp();
#pragma unroll
              for (int jq = 0; jq < 8; ++jq) {                    // row `lane`: elements n1 = 4 jq .. 4 jq + 3
                const float4 qr = *reinterpret_cast<const float4*>(&buf[lane * kRowFloats + 4 * jq]);
                const float4 qi = *reinterpret_cast<const float4*>(&buf[lane * kRowFloats + 32 + 4 * jq]);
                R[2 * jq] = make_float2(qr.x, qr.y); R[2 * jq + 1] = make_float2(qr.z, qr.w);
                I[2 * jq] = make_float2(qi.x, qi.y); I[2 * jq + 1] = make_float2(qi.z, qi.w);
              }
              if (MODE == MODE_GL_ITER && half == 0) {
                // the exchange buffer is idle until the inverse transform: stream this frame's |S| row into it now,
                // so that the HBM latency hides behind the second pass
                __syncwarp();
                s_off = row_to_smem_async(buf, a.spec + row * kF, a.spec, a.spec_end, lane);
              }
            }
          }
        }  // halves

        if constexpr (MODE == MODE_ANALYSIS) {
          // ---------------------------------------------------------------- spectrum out
          // X[k] = (E2 + G_k D2)/2, X[1024-k] = conj(E2 - G_k D2)/2 ; undo the circular shift for complex output
          float* magbuf = buf;
          float2* cout = reinterpret_cast<float2*>(a.cplx_out) + row * kF;
          __syncwarp();
          static_for<0, 8>([&](auto mc) {
            constexpr int m = decltype(mc)::value;
            constexpr int ms = (m == 0) ? 0 : 16 - m;
            const float s0r = l0 ? R[ms].x : R[15 - m].y, s0i = l0 ? I[ms].x : I[15 - m].y;
            const float s1r = l0 ? R[15 - m].y : R[15 - m].x, s1i = l0 ? I[15 - m].y : I[15 - m].x;
            const float2 BR = shfl2(s0r, s1r, partner), BI = shfl2(s0i, s1i, partner);
            const int k0 = 64 * m + lane;
            const float4 gq = g4[m * 32 + lane];
            const float2 GX = make_float2(gq.x, gq.y), GY = make_float2(gq.z, gq.w);
            const float2 E2R = __fadd2_rn(R[m], BR), E2I = __fadd2_rn(I[m], neg2(BI));
            const float2 D2R = __fadd2_rn(R[m], neg2(BR)), D2I = __fadd2_rn(I[m], BI);
            float2 XkR = __ffma2_rn(GX, D2R, E2R);
            XkR = __ffma2_rn(neg2(GY), D2I, XkR);
            float2 XkI = __ffma2_rn(GX, D2I, E2I);
            XkI = __ffma2_rn(GY, D2R, XkI);
            float2 XpR = __ffma2_rn(E2R, splat(2.0f), neg2(XkR));
            float2 XpI = __ffma2_rn(E2I, splat(-2.0f), XkI);
            XkR = __fmul2_rn(XkR, splat(0.5f)); XkI = __fmul2_rn(XkI, splat(0.5f));
            XpR = __fmul2_rn(XpR, splat(0.5f)); XpI = __fmul2_rn(XpI, splat(0.5f));
            if (m == 0 && l0) { XkI.x = 0.0f; XpI.x = 0.0f; }
            if constexpr (SRC == OUT_COMPLEX) {
              // frame tap m sits at transform sample lpad + m: X_true[k] = X[k] * exp(-j 2 pi k lpad / n_fft)
              cout[k0] = cmul(make_float2(XkR.x, XkI.x), shift_phasor(k0, lpad, -1.0f));
              cout[k0 + 32] = cmul(make_float2(XkR.y, XkI.y), shift_phasor(k0 + 32, lpad, -1.0f));
              cout[1024 - k0] = cmul(make_float2(XpR.x, XpI.x), shift_phasor(1024 - k0, lpad, -1.0f));
              cout[992 - k0] = cmul(make_float2(XpR.y, XpI.y), shift_phasor(992 - k0, lpad, -1.0f));
            } else {
              magbuf[k0] = sqrt_fast(XkR.x * XkR.x + XkI.x * XkI.x);
              magbuf[k0 + 32] = sqrt_fast(XkR.y * XkR.y + XkI.y * XkI.y);
              magbuf[1024 - k0] = sqrt_fast(XpR.x * XpR.x + XpI.x * XpI.x);
              magbuf[992 - k0] = sqrt_fast(XpR.y * XpR.y + XpI.y * XpI.y);
            }
          });
          if constexpr (SRC == OUT_COMPLEX) {
            if (l0) {   // k = 512: X = conj(Z[512])
              const float2 Xc = make_float2(R[8].x, -I[8].x);
              cout[512] = cmul(Xc, shift_phasor(512, lpad, -1.0f));
            }
          } else {
            // k = 512 is lane 0's: every lane stores lane 0's value.  A lane-dependent branch here is not reconverged before
            // the contraction below, which then runs once for lane 0 and once for the other 31 (measured in feat_stream.cuh)
            const float am = sqrt_fast(R[8].x * R[8].x + I[8].x * I[8].x);
            magbuf[512] = __shfl_sync(0xffffffffu, am, 0);
          }
          if constexpr (SRC == OUT_FEATURES) {
            __syncwarp();
            if (a.lin_out != nullptr) {
              float* out = a.lin_out + row * kF;
              for (int k = lane; k < kF; k += 32) out[k] = amp_to_norm_db(magbuf[k], g);
            }
            if (a.mel_out != nullptr) {
              float* out = a.mel_out + row * g.num_mels;
              if (g.mel_smem_floats > 0) {
                // compact basis in shared memory: (first tap, first bin) per filter, then the taps
                const int2* mdesc = reinterpret_cast<const int2*>(smem + ly.sm_total);
                const float* mval = smem + ly.sm_total + 2 * (g.num_mels + 1);
                for (int m = lane; m < g.num_mels; m += 32) {
                  const int2 d0 = mdesc[m];
                  const int cnt = mdesc[m + 1].x - d0.x;
                  const float* mv = mval + d0.x;
                  const float* mg = magbuf + d0.y;
                  float acc = 0.0f;
                  for (int c = 0; c < cnt; ++c) acc = fmaf(mv[c], mg[c], acc);
                  out[m] = amp_to_norm_db(acc, g);
                }
              } else {
                for (int m = lane; m < g.num_mels; m += 32) {
                  const int lo = tb.mel_lo[m], cnt = tb.mel_cnt[m];
                  const float* mv = tb.mel_val + m * tb.mel_ld;
                  float acc = 0.0f;
                  for (int c = 0; c < cnt; ++c) acc = fmaf(__ldg(mv + c), magbuf[lo + c], acc);
                  out[m] = amp_to_norm_db(acc, g);
                }
              }
            }
          }
        } else {
          // ---------------------------------------------------------------- the warp's overlap-add slot
          // element n2 = conj(z'[lane + 32 n2]):  sample 2q = Re, sample 2q+1 = -Im  (q = lane + 32 n2); the slot keeps
          // the even samples and the (un-negated) odd samples in two planes; window and sign are applied in the
          // overlap-add
          __syncwarp();                       // all lanes done reading the exchange buffer: it becomes the slot
#pragma unroll
          for (int m = 0; m < 16; ++m) {
            if (2 * m < NZ) {
              const int q = lane + 64 * m;
              if (q < ly.half) { buf[q] = R[m].x; buf[kSlotPlane + q] = I[m].x; }
              if (q + 32 < ly.half) { buf[q + 32] = R[m].y; buf[kSlotPlane + q + 32] = I[m].y; }
            }
          }
        }
      }

      // next span: its loads are issued by each warp as soon as its frame is done, so they are in flight while the
      // warp waits for the others and while the tile is overlap-added (the planes are rewritten after that)
      const bool have_next = jt + 1 < jb;
      if constexpr (MODE != MODE_SYNTH) {
        if (have_next && !TTSA_SKIP(a, 2)) stage_load(jt + 1);
      }
      __syncthreads();                             // slots complete; planes consumed

      if constexpr (MODE != MODE_ANALYSIS) {
        // -------------------------------------------------------------------- overlap-add + window + 1/(N wss) + store
        // each thread owns a residue rr (mod hop): acc[j] is span sample j*hop + rr; frame f adds its taps
        // rr + d*hop (d < ND), weighted by the synthesis window, to acc[f + d].  All register indices are static;
        // lanes read consecutive addresses of alternating slot planes.  The carry (samples that later frames still
        // add to) is read and re-written by the same thread.
        const int fv_lo = first_needed > t0 ? first_needed - t0 : 0;
        const int fv_hi = (sfb - t0) < kNF ? (sfb - t0) : kNF;
        // the utterance's last tile also flushes what would be its carry (samples up to hop*(T-1) end there); a fine
        // segment writes [own_lo, out_len): what lies before belongs to the previous segment, what follows to the next
        const int out_len = FINE ? (sfb >= T ? ly.span_len : (sfb - t0) * ly.hop) : ((t0 + kNF >= T) ? ly.span_len : kNF * ly.hop);
        const int own_lo = FINE ? (sfa - t0) * ly.hop : 0;
        float* __restrict__ dst = a.wav_out + woff;
        const bool interior = !FINE && write_out && fv_lo == 0 && fv_hi == kNF && i0 >= 0 && t0 + kNF < T &&
                              i0 + kNF * ly.hop <= L && t0 >= ND - 1;
        // a full fine tile away from the utterance's ends: positions [nwarm*hop, 8*hop) are written, nothing is carried
        const bool interior_fine = FINE && fs > 0 && fv_hi == kNF && sfb < T && i0 >= 0 && i0 + kNF * ly.hop <= L &&
                                   t0 >= ND - 1;
        // finished sample `val` at span position sidx = j*hop + rr: scale and store, or keep as carry
        auto emit = [&](int j, int rr, int sidx, float val) {
          if (sidx >= ly.span_len) return;
          if (FINE && (sidx < own_lo || sidx >= out_len)) return;
          if (sidx < out_len) {
            const int i = i0 + sidx;
            if (write_out && i >= 0 && i < L) {
              float inv = pw[rr];
              if (t0 + j - (ly.win - 1 - rr) / ly.hop < 0 || t0 + j > T - 1) {   // some overlapping frame does not exist
                float ws = 0.0f;
                for (int d = 0, m = rr; m < ly.win; ++d, m += ly.hop) {
                  const int tt = t0 + j - d;
                  if (tt >= 0 && tt < T) ws = fmaf(wsyn[m], wsyn[m], ws);
                }
                inv = ws > 1.17549435e-38f ? kInvN / ws : kInvN;              // librosa: divide only where wss > tiny
              }
              dst[i] = val * inv;
            }
          } else {
            carry[sidx - out_len] = val;
          }
        };
        // Residue ownership: warps 0..3 take the even residues 2i, warps 4..7 the odd residues 2i+1 (i = tid mod 128).
        // Tap m = rr + d*hop of residue rr = 2i + par sits at slot offset i + c(par, d), c = plane(par + d*hop) +
        // (par + d*hop)/2: a per-thread base plus constants that fold into the load's immediate when the geometry
        // is fixed at compile time; lanes read consecutive words.
        auto ola = [&](auto parc, auto rb) {
          constexpr int par = decltype(parc)::value;
          const int i2 = (tid & (kThreads / 2 - 1)) + rb;
          const int rr = 2 * i2 + par;
          if (rr >= ly.hop) return;
          const float* const sl0 = smem + i2;
          float acc[kNF + ND - 1];
          float wreg[ND];
          int off[ND];
#pragma unroll
          for (int d = 0; d < ND; ++d) {
            const int m = rr + d * ly.hop;                // window tap; odd taps carry the conjugation sign
            const int e = par + d * ly.hop;
            wreg[d] = m < ly.win ? wsyn[m] : 0.0f;
            off[d] = m < ly.win ? ((e & 1) ? kSlotPlane : 0) + (e >> 1) : 0;
          }
#pragma unroll
          for (int j = 0; j < kNF + ND - 1; ++j) {
            const int sidx = j * ly.hop + rr;
            acc[j] = (has_carry && j < ND - 1 && sidx < ly.carry_len) ? carry[sidx] : 0.0f;
          }
          if (interior) {
#pragma unroll
            for (int f = 0; f < kNF; ++f) {
#pragma unroll
              for (int d = 0; d < ND; ++d) acc[f + d] = fmaf(sl0[f * kBufFloats + off[d]], wreg[d], acc[f + d]);
            }
            const float inv = pw[rr];
            float* __restrict__ po = dst + (i0 + rr);
#pragma unroll
            for (int j = 0; j < kNF; ++j) po[j * ly.hop] = acc[j] * inv;
#pragma unroll
            for (int j = kNF; j < kNF + ND - 1; ++j) {
              const int c = (j - kNF) * ly.hop + rr;
              if (c < ly.carry_len) carry[c] = acc[j];
            }
          } else if (interior_fine) {
#pragma unroll
            for (int f = 0; f < kNF; ++f) {
#pragma unroll
              for (int d = 0; d < ND; ++d) acc[f + d] = fmaf(sl0[f * kBufFloats + off[d]], wreg[d], acc[f + d]);
            }
            const float inv = pw[rr];
            float* __restrict__ po = dst + (i0 + rr);
#pragma unroll
            for (int j = 0; j < kNF; ++j)
              if (j >= ly.nwarm) po[j * ly.hop] = acc[j] * inv;
          } else {
#pragma unroll
            for (int f = 0; f < kNF; ++f) {
              if (f >= fv_lo && f < fv_hi) {
#pragma unroll
                for (int d = 0; d < ND; ++d) acc[f + d] = fmaf(sl0[f * kBufFloats + off[d]], wreg[d], acc[f + d]);
              }
            }
#pragma unroll 1
            for (int j = 0; j < kNF + ND - 1; ++j) {
              float val = 0.0f;
#pragma unroll
              for (int jj = 0; jj < kNF + ND - 1; ++jj) val = (jj == j) ? acc[jj] : val;
              emit(j, rr, j * ly.hop + rr, val);
            }
          }
        };
        if (!TTSA_SKIP(a, 2)) {
          if (tid < kThreads / 2) {
            ola(IntC<0>{}, IntC<0>{});
          } else {
            ola(IntC<1>{}, IntC<0>{});
          }
        }
        // residues beyond the thread count (hop 275 = 256 + 19): one small item per (residue, j mod kNF) spread over
        // all warps; a second full round for the few threads that own one would sit on the tile's critical path
        // (measured: 0.160 vs 0.155 ms per iteration)
        if (ly.hop > kThreads && !TTSA_SKIP(a, 2)) {
          const int nl = ly.hop - kThreads;
          for (int it = tid; it < nl * kNF; it += kThreads) {
            const int jl = it / nl, rr = kThreads + it - jl * nl;
            const int sa = jl * ly.hop + rr, sb = sa + kNF * ly.hop;
            float va = (has_carry && sa < ly.carry_len) ? carry[sa] : 0.0f;
            float vb = 0.0f;
            for (int d = 0, m = rr; d < ND && m < ly.win; ++d, m += ly.hop) {
              const int po = ((m & 1) ? kSlotPlane : 0) + (m >> 1);
              const float w = wsyn[m];
              const int fa = jl - d, fb = fa + kNF;
              if (fa >= fv_lo && fa < fv_hi) va = fmaf(smem[fa * kBufFloats + po], w, va);
              if (fb >= fv_lo && fb < fv_hi) vb = fmaf(smem[fb * kBufFloats + po], w, vb);
            }
            if (interior) {   // owned, inside the signal, every overlapping frame exists: no edge handling
              dst[i0 + sa] = va * pw[rr];
              if (jl < ND - 1 && sb < ly.span_len) carry[sb - kNF * ly.hop] = vb;
            } else if (interior_fine) {
              if (jl >= ly.nwarm) dst[i0 + sa] = va * pw[rr];
            } else {
              emit(jl, rr, sa, va);
              if (jl < ND - 1) emit(jl + kNF, rr, sb, vb);
            }
          }
        }
        has_carry = true;
      }

      if constexpr (MODE != MODE_SYNTH) {
        if (have_next && !TTSA_SKIP(a, 2)) stage_store(jt + 1);
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
