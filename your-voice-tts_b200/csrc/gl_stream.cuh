// Griffin-Lim iteration, "warp stream" form (utils/audio.py:186-188): every warp owns a private run of CONSECUTIVE
// frames of one utterance and carries the overlap-add of its run in a private shared-memory ring, so the steady state
// has no CTA-wide barrier, no shared staging buffer and no separate overlap-add phase:
//
//   per frame (one warp):  cp.async'ed input span -> window -> FFT (32 x 32, one exchange) -> |S| e^{j angle X}
//                          -> inverse FFT (same code, conjugate trick) -> window, ONE pass over the ring:
//                             the hop samples no later frame touches get their last term, * 1/(N wss) -> global;
//                             the rest is accumulated in place; the part of the window no earlier frame reached is
//                             stored without a load (so the ring never needs zeroing)
//
// The 16 warps of the one CTA per SM drift apart and sit in different phases (FP32-bound transform passes,
// shared-memory-bound exchanges, latency-bound per-bin step), which is what keeps the pipes busy; frame_kernel<GL_ITER>
// (frame_kernels.cuh) runs its 8 warps in lock step between two CTA barriers per tile and spends 25 % of its warp
// time in the overlap-add / staging / barrier phases this kernel does not have.
//
// Alignment.  A frame starts at sample t*hop - win/2 of its utterance, which is odd for every other frame when the hop
// is odd (275).  All shared and global accesses here are 8-byte (sample pair) accesses on EVEN absolute sample
// positions: an odd frame is processed as the frame that starts one sample earlier with the window shifted by one tap
// (w'[0] = 0).  A shift of the whole frame is a phase ramp, which the per-bin projection S X/|X| and the inverse
// transform undo exactly (the same argument as the window-relative coordinates of frame_kernels.cuh).
//
// Run boundaries.  The host cuts the flattened frame list into one contiguous range per warp (ttsa_batch, wps_cut).
// A range that starts in the middle of an utterance lacks the contributions of the `kWarm` frames before it: its first
// win - hop samples (the "head zone") are stored as RAW partial sums and the warp then publishes a flag in global memory
// (release); the warp that owns the frames before the cut -- in this CTA or in the previous one -- keeps the matching
// partial sums in its ring and, at the very end of its run, waits for that flag (acquire), adds the two halves and
// applies 1/(N wss).  Nothing is recomputed, every output sample is still written last by a fixed thread with a fixed
// summation order (deterministic), and the only wait is on a warp that published its zone a whole run earlier.  The
// waiting warp depends on a warp that itself waits for nobody before publishing, so CTAs need not be co-resident.
#pragma once
#include "frame_kernels.cuh"

namespace ttsa {

constexpr int kWpsThreads = kWpsWarps * 32;

__device__ __forceinline__ float2 ld_volatile_f2(const float* p) {        // another SM wrote it: read through L2
  float2 v;
  asm volatile("ld.relaxed.gpu.global.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu(int* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// Timing probes (experiment builds only, -DTTSA_PROBE=mask: results are wrong on purpose): 1 no 32-point passes,
// 2 no inter-pass twiddles, 4 no exchange, 8 no per-bin step, 16 no ring accumulation, 32 no emission, 64 no input load, 128 no asynchronous copies (|S| rows, input spans)
#ifndef TTSA_PROBE
#define TTSA_PROBE 0
#endif
#ifndef TTSA_RING_GROUP
#define TTSA_RING_GROUP 8
#endif
#ifndef TTSA_ZONE_DEFER
#define TTSA_ZONE_DEFER 1
#endif
constexpr int kProbe = TTSA_PROBE;

// One row of the spectrogram / one input span into the warp's buffer by ONE bulk asynchronous copy (the enclosing
// 16-byte-aligned range; completion on the warp's mbarrier) instead of nine 16-byte cp.async per lane: no address
// arithmetic, no load/store-unit wavefronts, one issuing lane.  Ranges that would leave the tensor fall back to cp.async.
// Returns the landing offset of element 0 (0..3); `bulk` tells which completion mechanism to wait on.
template <int N>
__device__ __forceinline__ int span_to_smem_bulk_n(float* dst, const float* ptr, const float* lo, const float* hi, int lane,
                                                   unsigned mbar, bool& bulk) {
  const int off = (int)((reinterpret_cast<unsigned long long>(ptr) >> 2) & 3ull);
  const float* base = ptr - off;                      // 16-byte aligned
  constexpr int kMaxChunks = (3 + N + 3) >> 2;
  bulk = base >= lo && base + 4 * kMaxChunks <= hi;
  if (bulk) {
    if (lane == 0) {
      const unsigned bytes = (unsigned)((off + N + 3) >> 2) << 4;
      const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(d), "l"(base), "r"(bytes), "r"(mbar) : "memory");
    }
    return off;
  }
  return span_to_smem_async(dst, ptr, N, lo, hi, lane);
}
// An input span that lies inside its utterance's slot of the packed waveform buffer (slots start 16-byte aligned and are
// padded to whole 16-byte chunks, so the enclosing aligned range is inside the slot as well): always one bulk copy.
template <int N>
__device__ __forceinline__ int span_to_smem_bulk_inside(float* dst, const float* ptr, int lane, unsigned mbar) {
  const int off = (int)((reinterpret_cast<unsigned long long>(ptr) >> 2) & 3ull);
  if (lane == 0) {
    const unsigned bytes = (unsigned)((off + N + 3) >> 2) << 4;
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(d), "l"(ptr - off), "r"(bytes), "r"(mbar) : "memory");
  }
  return off;
}
__device__ __forceinline__ float ld_relaxed_gpu_f32(const float* p) {      // coherent at gpu scope (not served by a stale L1 line)
  float v;
  asm volatile("ld.relaxed.gpu.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void mbar_wait(unsigned mbar, unsigned parity) {
  unsigned done = 0;
  while (!done) {
    asm volatile("{\n\t.reg .pred q;\n\tmbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2;\n\tselp.u32 %0, 1, 0, q;\n\t}\n"
                 : "=r"(done) : "r"(mbar), "r"(parity) : "memory");
  }
}

__device__ __forceinline__ int ld_relaxed_gpu_s32(const int* p) {
  int v;
  asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// Wait until *p >= v (flags only grow).  One lane polls with an acquire load (every acquire invalidates the SM's L1, so
// not 32 of them per poll); the warp barrier extends the ordering to the other lanes, whose reads of the published data
// are L2 loads (ld.relaxed.gpu) or bulk copies issued by the polling lane.
__device__ __forceinline__ void wait_flag_ge(const int* p, int v, bool poller) {
  if (poller) while (ld_acquire_gpu(p) < v) __nanosleep(64);
  __syncwarp();
}
// Publish flag *p = v after everything the warp wrote: one fence per lane, then a relaxed store by one lane (fence +
// relaxed store is the release pattern; st.release would add a second fence, __threadfence() a sequentially consistent one).
__device__ __forceinline__ void publish_flag(int* p, int v, bool writer) {
  asm volatile("fence.acq_rel.gpu;" ::: "memory");
  __syncwarp();
  if (writer) asm volatile("st.relaxed.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

template <int HOP, int WIN>
struct WpsGeo {
  static_assert(WIN % 2 == 0 && HOP >= 2 && WIN >= HOP, "even window not shorter than the hop");
  static constexpr int kOdd = HOP & 1;
  static constexpr int kNP = WIN / 2 + kOdd;           // sample pairs of an (aligned) frame
  static constexpr int kRows = (kNP + 31) / 32;        // rows of 32 pairs
  static constexpr int kMH = (kRows + 1) / 2;          // packed register pairs that carry data
  static constexpr int kRH = 32 * kRows;               // ring length in pairs
  static constexpr int kWarm = (WIN - 1) / HOP;        // frames before t that overlap frame t's first sample
  static constexpr int kZone = WIN - HOP;              // samples of a run's first frames that earlier frames also add to
  static constexpr int kEmitIters = ((HOP + 1) / 2 + 31) / 32;
  static constexpr int kFlushIters = (kNP + 31) / 32;
  static_assert(kRows <= 20, "window too long for this kernel");
  static constexpr int kCntMin = HOP / 2, kCntMax = (HOP + 1) / 2;   // pairs that leave the ring per interior frame
  static constexpr int kEmitRows = (kCntMax + 31) / 32;
  // utterance edges: samples [0, kHead) lack frames before the first one, the last kTail samples the frame after the
  // last one; their 1 / (n_fft wss) comes from two plan tables when the utterance is long enough for both to be exact
  static constexpr int kHead = kWarm * HOP - WIN / 2 > 0 ? kWarm * HOP - WIN / 2 : 0;
  static constexpr int kTail = WIN / 2 - HOP > 0 ? WIN / 2 - HOP : 0;
  static constexpr int kEdgeMinT = 2 * kWarm + 2;
  static constexpr int kPwx = (64 * kEmitRows + 4 + 3) / 4 * 4;  // entries of the shifted 1/(N wss) table
  static_assert(kPwx >= HOP + 2, "hop too long for the emission table");
  // shared memory (floats): per-warp buffers, then the constant tables as ONE host-built image (Tables::wps_image)
  static constexpr int kWarpFloats = kBufFloats + 2 * kRH;
  static constexpr int sm_img = kWpsWarps * kWarpFloats;
  static constexpr int sm_tw = sm_img;                           // float4[16][32]
  static constexpr int sm_g = sm_tw + 2048;                      // float4[8][32]
  static constexpr int kWS = kWpsWinStride;                      // window taps per lane: pair q = lane + 32 n at [lane][n]
  static_assert(kRows <= kWS && (kWS / 4) % 2 == 1, "lane rows of the window tables: 16-byte loads without bank conflicts");
  static constexpr int sm_wA = sm_g + 1024;                      // [32][kWS]  w[2q]
  static constexpr int sm_wB = sm_wA + 32 * kWS;                 // [33][kWS]  row lane + 1: w[2q+1]; row lane: w[2(q-1)+1] = w[2q-1]
  static constexpr int sm_pwx = sm_wB + 33 * kWS;                // [kPwx]     pwx[j] = 1 / (n_fft wss[(j - 1) mod HOP])
  static constexpr int image_floats = 2048 + 1024 + 65 * kWS + kPwx;
  static constexpr int sm_mbar = sm_img + image_floats;          // 8-byte aligned (image_floats is a multiple of 4)
  static constexpr int sm_total = sm_mbar + 4 + 4 * kWpsWarps;   // + two mbarriers per warp (|S| row, input span)
  static constexpr bool kFits = sm_total * 4 <= 227 * 1024;
};

// -DTTSA_WPS_TRACE (experiment builds, tools/wps_trace.py): per-warp global-timer stamps of every iteration's phases
#ifdef TTSA_WPS_TRACE
#define WPS_STAMP(k) do { if (l0) { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); \
    a.wps_trace[((size_t)wi * 64 + (it & 63)) * 8 + (k)] = t_; } } while (0)
#else
#define WPS_STAMP(k) do {} while (0)
#endif

template <int SRC, bool SC, int HOP, int WIN, bool FUSE>
__global__ void __launch_bounds__(kWpsThreads, 1)
gl_stream_kernel(const Geo g, const Tables tb, const BatchDev bd, const WpsDev wp, const FrameArgs a) {
  using G = WpsGeo<HOP, WIN>;
  extern __shared__ __align__(16) float smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = warp_index();
  constexpr float kInvN = 1.0f / (float)kNfft;
  constexpr float kTiny = 1e-37f;
  constexpr float kPhaseEps = 1e-18f;

  float* const buf = smem + warp * G::kWarpFloats;                 // exchange buffer / |S| row / input span landing zone
  float2* const ring = reinterpret_cast<float2*>(buf + kBufFloats);
  const float4* const tw4 = reinterpret_cast<const float4*>(smem + G::sm_tw);
  const float4* const g4 = reinterpret_cast<const float4*>(smem + G::sm_g);
  const float* const wA = smem + G::sm_wA;
  const float* const wB = smem + G::sm_wB;
  const float* const pwx = smem + G::sm_pwx;

  // 1 / (n_fft * window sum of squares) at sample i of an utterance with T frames, over the frames that exist
  // (librosa: divide only where wss > tiny); the interior uses the periodic table pwx instead
  auto inv_wss = [&](int i, int T) {
    float ws = 0.0f;
    const int tq = (i + WIN / 2) / HOP;                            // last frame whose window starts at or before i
    for (int tt = tq; tt >= 0 && tt > tq - (G::kWarm + 1); --tt) {
      const int mtap = i - (tt * HOP - WIN / 2);
      if (tt < T && mtap >= 0 && mtap < WIN) {
        const int q = mtap >> 1;
        const float wv = (mtap & 1) ? wB[((q & 31) + 1) * G::kWS + (q >> 5)] : wA[(q & 31) * G::kWS + (q >> 5)];
        ws = fmaf(wv, wv, ws);
      }
    }
    return ws > 1.17549435e-38f ? kInvN / ws : kInvN;
  };
  // the same through the plan's edge tables (Tables::edge_head / edge_tail) and the periodic table
  auto edge_inv = [&](int i, int T, int L) {
    if (T < G::kEdgeMinT) return inv_wss(i, T);
    if (i < G::kHead) return __ldg(tb.edge_head + i);
    if (i >= L - G::kTail) return __ldg(tb.edge_tail + (i - (L - G::kTail)));
    return pwx[(i + WIN / 2) % HOP + 1];
  };

  // ---- prologue (independent of the previous kernel's output): the table image by one bulk copy, this warp's range
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const unsigned mbar = (unsigned)__cvta_generic_to_shared(smem + G::sm_mbar);
  if (tid == 0) {
    const unsigned dst = (unsigned)__cvta_generic_to_shared(smem + G::sm_img);
    constexpr unsigned bytes = (unsigned)G::image_floats * 4u;
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar) : "memory");
    for (int i = 0; i < 2 * kWpsWarps; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar + 16 + 8 * i) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(tb.wps_image), "r"(bytes), "r"(mbar) : "memory");
  }
  const int wi = blockIdx.x * kWpsWarps + warp;
  const int fa = uni(wp.cut[wi]), fb = uni(wp.cut[wi + 1]);
  __syncthreads();                                                 // the mbarrier is initialised for everyone
  {
    unsigned done = 0;
    while (!done) {
      asm volatile("{\n\t.reg .pred q;\n\tmbarrier.try_wait.parity.shared::cta.b64 q, [%1], 0;\n\tselp.u32 %0, 1, 0, q;\n\t}\n"
                   : "=r"(done) : "r"(mbar) : "memory");
    }
  }
  // programmatic dependent launch: the previous kernel's waveform is complete from here on
  asm volatile("griddepcontrol.wait;" ::: "memory");

  const unsigned mbar_s = mbar + 16 + 16 * warp, mbar_x = mbar_s + 8;   // this warp's copy barriers and their phases
  unsigned ph_s = 0, ph_x = 0;
  const int partner = (32 - lane) & 31;
  const bool l0 = lane == 0;
  // ---- iterations of this launch.  Iteration `it` reads the waveform iteration it - 1 wrote (ping-pong between wav_in and
  // wav_out).  A kernel boundary between iterations is NOT needed: a warp's input spans cover its own output region plus
  // the head zone of the next run (which it finishes itself), and the only samples somebody else writes into its region are
  // its head zone, finished by the owner of the frames before the cut at the end of THAT warp's run.  So a warp whose range
  // starts inside an utterance waits for the previous warp's "iteration done" flag (release / acquire at gpu scope) and
  // every other warp starts at once.  The same wait orders the write-after-read on the ping-pong buffers (the previous
  // warp's last frames read into this warp's region).  SMs therefore drift apart by whole frames between iterations: a
  // 14-frame run no longer makes 2 367 other warps wait at a kernel boundary, and launch / prologue / drain are paid once.
  const int n_it = FUSE ? a.wps_iters : 1;                        // (the loop costs the one-iteration kernel 2.7 % when it is not compiled out)
#pragma unroll 1
  for (int it = 0; it < n_it; ++it) {
  const int epoch = a.wps_epoch + it;
  int u = uni(wp.u0[wi]);                                            // utterance of frame fa (host-built)
  WPS_STAMP(0);
  if (it > 0 && fa < fb) {
    if (wp.tsum[u] < fa) {                                         // the range starts inside utterance u
      int pw = wi - 1;
      while (wp.cut[pw + 1] == wp.cut[pw]) --pw;                   // some earlier warp owns frame fa - 1
      wait_flag_ge(a.wps_done + pw, epoch - 1, l0);
    }
    // the bulk copies below read (async proxy) what ordinary stores of the previous iteration wrote
    asm volatile("fence.proxy.async.global;" ::: "memory");
  }
  WPS_STAMP(1);
  const float* const wav_rd = (it & 1) ? a.wav_out : a.wav_in;
  float* const wav_wr = (it & 1) ? const_cast<float*>(a.wav_in) : a.wav_out;
  int f = fa;
  while (f < fb) {
    while (wp.tsum[u + 1] <= f) ++u;                               // skips empty utterances
    const int tsu = uni(wp.tsum[u]);
    const int T = uni(bd.T[u]);
    const int t_begin = f - tsu;
    const int t_end = min(T, fb - tsu);
    f = tsu + t_end;
    const int L = uni(bd.wav_len[u]);
    if (L <= 0) continue;
    const long long woff = uni(bd.wav_off[u]);
    const float* __restrict__ src = wav_rd + woff;
    float* __restrict__ dst = wav_wr + woff;
    const float* spec_row0 = a.spec + uni(bd.frame_off[u]) * kF;

    // a run that starts inside the utterance leaves its first kZone samples as raw partial sums; the owner of the
    // frames before the cut finishes them at the end of its run
    const int zone_end = t_begin > 0 ? (t_begin - 1) * HOP + WIN / 2 : -(1 << 30);
    int zone_state = zone_end > 0 ? 1 : 0;                         // 1: head zone not stored yet, 2: stored, flag not published
    float sc_num = 0.0f, sc_den = 0.0f;

    // input span of frame t: samples [a0, a0 + WIN + p) of the utterance, a0 even; asynchronous when it needs no reflection
    auto span_fast = [&](int t) {
      const int s0 = t * HOP - WIN / 2;
      const int a0 = s0 - (s0 & 1);
      return a0 >= 0 && a0 + WIN + 2 <= L;
    };
    const bool edge_async = L >= 2 * (WIN + 2);                    // one fold of the reflection per span, source inside the copy
    // a frame whose span was copied by the edge path: fill the samples outside [0, L) from their mirror images
    auto fill_reflected = [&](int a0, int x_off) {
      const int A = a0 - x_off;
      const int c_lo = max(A, 0), c_hi = min(a0 + WIN + 2, L);
      // edge_async: one fold.  Sample idx < 0 mirrors to -idx, idx >= L to 2 (L - 1) - idx; a mirror image outside the copy
      // belongs to a zero-tap padding element of the aligned frame
#pragma unroll 1
      for (int idx = a0 + lane; idx < 0; idx += 32) buf[idx - A] = (-idx < c_hi) ? buf[-idx - A] : 0.0f;
#pragma unroll 1
      for (int idx = L + lane; idx < a0 + WIN + 2; idx += 32) {
        const int r = 2 * (L - 1) - idx;
        buf[idx - A] = (r >= c_lo) ? buf[r - A] : 0.0f;
      }
      __syncwarp();
    };
    auto span_issue = [&](int t) {                                 // returns the landing offset (0 or 2)
      const int s0 = t * HOP - WIN / 2;
      const int a0 = s0 - (s0 & 1);
      if (span_fast(t)) return span_to_smem_bulk_inside<WIN + 2>(buf, src + a0, lane, mbar_x);
      if (edge_async) {
        // the first and last frames of an utterance (np.pad(..., mode='reflect')): the part of the span that exists, from 8
        // samples before the frame (a reflected tap of the last frame can point just below it), by one bulk copy; the
        // reflected part is filled in from shared memory when the frame is consumed (fill_reflected)
        const int A = (a0 - 8) & ~3;                               // sample that lands at buf[0]
        if (lane == 0) {
          const int c_lo = max(A, 0), c_hi = (min(a0 + WIN + 2, L) + 3) & ~3;
          const unsigned bytes = (unsigned)(c_hi - c_lo) << 2;
          const unsigned d = (unsigned)__cvta_generic_to_shared(buf + (c_lo - A));
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar_x), "r"(bytes) : "memory");
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                       ::"r"(d), "l"(src + c_lo), "r"(bytes), "r"(mbar_x) : "memory");
        }
        return a0 - A;
      }
      // short utterances (several folds of the reflection inside one span).  Loads at gpu scope: with
      // several iterations per launch the same address is re-read after other SMs rewrote it, so no L1-allocating copy
#pragma unroll 1
      for (int m0 = lane; m0 < WIN + 2; m0 += 32 * 6) {
        float v[6];
#pragma unroll
        for (int k = 0; k < 6; ++k) { const int m = m0 + 32 * k; v[k] = m < WIN + 2 ? ld_relaxed_gpu_f32(src + reflect_index(a0 + m, L)) : 0.0f; }
#pragma unroll
        for (int k = 0; k < 6; ++k) { const int m = m0 + 32 * k; if (m < WIN + 2) buf[m] = v[k]; }
      }
      return 0;
    };
    int x_off = 0;
    if (!(kProbe & (128 | 1024))) x_off = span_issue(t_begin);
    bool s_bulk = false;
    int base = (((t_begin * HOP - WIN / 2) >> 1) + WIN) % G::kRH;  // ring slot of the first frame's first pair (a0 / 2, made positive)

    int newq = 0;                                                  // first pair of the frame that no earlier frame of this run reached
#pragma unroll 1
    for (int t = t_begin; t < t_end; ++t) {
      const int s0 = t * HOP - WIN / 2;
      const int p = s0 & 1;
      const int a0 = s0 - p;
      // window of the even / odd sample of pair q = lane + 32 n: pe[n], po[n] (an odd frame starts one sample early)
      const float* const pe = (p ? wB : wA) + lane * G::kWS;
      const float* const po = (p ? wA : wB + G::kWS) + lane * G::kWS;

      float2 R[16], I[16];
      int s_off = 0;
#pragma unroll 1
      for (int half = 0; half < 2; ++half) {
        if (half == 0) {
          // ---------------------------------------------------------------- input span -> windowed packed frame
          if (kProbe & 2048) {
          } else if (edge_async || span_fast(t)) { mbar_wait(mbar_x, ph_x); ph_x ^= 1u; }   // the span came by a bulk copy
          else cp_async_wait_all();
          __syncwarp();
          if (edge_async && !span_fast(t)) fill_reflected(a0, x_off);
          const float2* const xp = reinterpret_cast<const float2*>(buf + x_off) + lane;
#pragma unroll
          for (int m = 0; m < 16; ++m) {
            if (m < G::kMH && !(kProbe & 64)) {
              // pairs past the window hold stale exchange data: force zeros (their window taps are zero, but 0 * NaN of
              // another utterance's bad spectrum must not leak into this one)
              float2 xa = xp[64 * m];
              if (64 * m + 31 >= WIN / 2 && lane + 64 * m >= WIN / 2 + p) xa = make_float2(0.0f, 0.0f);
              // taps of rows 2m, 2m + 1 by one 8-byte load per table (adjacent ones merge into 16-byte loads)
              float2 xb = make_float2(0.0f, 0.0f);
              const float2 we = (kProbe & 32768) ? make_float2(0.5f + 1e-3f * m, 0.3f) : *reinterpret_cast<const float2*>(pe + 2 * m);
              const float2 wo = (kProbe & 32768) ? make_float2(0.4f, 0.2f + 1e-3f * m) : *reinterpret_cast<const float2*>(po + 2 * m);
              if (2 * m + 1 < G::kRows) {
                xb = xp[64 * m + 32];
                if (64 * m + 63 >= WIN / 2 && lane + 64 * m + 32 >= WIN / 2 + p) xb = make_float2(0.0f, 0.0f);
              }
              R[m] = __fmul2_rn(make_float2(xa.x, xb.x), we);
              I[m] = __fmul2_rn(make_float2(xa.y, xb.y), wo);
            } else {
              R[m] = make_float2(0.0f, 0.0f);
              I[m] = make_float2(0.0f, 0.0f);
            }
          }
          __syncwarp();                          // the landing zone becomes the exchange buffer
        } else if (!(kProbe & 8)) {
          // ---------------------------------------------------------------- per-bin step -> conj(Z')   (as frame_kernel)
          float2 BR[8], BI[8];
          static_for<0, 8>([&](auto mc) {
            constexpr int m = decltype(mc)::value;
            constexpr int ms = (m == 0) ? 0 : 16 - m;
            const float s0r = l0 ? R[ms].x : R[15 - m].y, s0i = l0 ? I[ms].x : I[15 - m].y;
            const float s1r = l0 ? R[15 - m].y : R[15 - m].x, s1i = l0 ? I[15 - m].y : I[15 - m].x;
            BR[m] = shfl2(s0r, s1r, partner);
            BI[m] = shfl2(s0i, s1i, partner);
          });
          if (kProbe & 2048) {
          } else if (s_bulk) { mbar_wait(mbar_s, ph_s); ph_s ^= 1u; }   // this frame's |S| row (issued between the two forward passes)
          else cp_async_wait_all();
          __syncwarp();
          const float* srow = buf + s_off;
          float2 SR[8], SI[8];
          float2 z512 = make_float2(0.0f, 0.0f);
          static_for<0, 8>([&](auto mc) {
            constexpr int m = decltype(mc)::value;
            const int k0 = 64 * m + lane;
            const float4 gq = (kProbe & 16384) ? make_float4(0.7f, 0.6f, 0.5f + 1e-3f * m, 0.4f) : g4[m * 32 + lane];
            const float2 GX = make_float2(gq.x, gq.y), GY = make_float2(gq.z, gq.w);
            const float2 Sk = make_float2(spec_to_mag<SRC>(srow[k0], g), spec_to_mag<SRC>(srow[k0 + 32], g));
            const float2 Sp = make_float2(spec_to_mag<SRC>(srow[1024 - k0], g), spec_to_mag<SRC>(srow[992 - k0], g));
            const float2 E2R = __fadd2_rn(R[m], BR[m]), E2I = __fadd2_rn(I[m], neg2(BI[m]));
            const float2 D2R = __fadd2_rn(R[m], neg2(BR[m])), D2I = __fadd2_rn(I[m], BI[m]);
            float2 XkR = __ffma2_rn(GX, D2R, E2R);
            XkR = __ffma2_rn(neg2(GY), D2I, XkR);
            float2 XkI = __ffma2_rn(GX, D2I, E2I);
            XkI = __ffma2_rn(GY, D2R, XkI);
            float2 XpR = __ffma2_rn(E2R, splat(2.0f), neg2(XkR));
            const float2 XpI = __ffma2_rn(E2I, splat(-2.0f), XkI);
            XkR = __fadd2_rn(XkR, splat(kPhaseEps));               // np.angle(0) = 0 without a select (see frame_kernels.cuh)
            XpR = __fadd2_rn(XpR, splat(kPhaseEps));
            const float2 mk = __ffma2_rn(XkI, XkI, __ffma2_rn(XkR, XkR, splat(kTiny)));
            const float2 mp = __ffma2_rn(XpI, XpI, __ffma2_rn(XpR, XpR, splat(kTiny)));
            const float2 ik = make_float2(rsqrt_fast(mk.x), rsqrt_fast(mk.y));
            const float2 ip = make_float2(rsqrt_fast(mp.x), rsqrt_fast(mp.y));
            const float2 fk = __fmul2_rn(Sk, ik), fp = __fmul2_rn(Sp, ip);
            const float2 YkR = __fmul2_rn(XkR, fk), YkI = __fmul2_rn(XkI, fk);
            const float2 YpR = __fmul2_rn(XpR, fp), YpI = __fmul2_rn(XpI, fp);
            if (SC) {
              const float2 dk = __ffma2_rn(__fmul2_rn(mk, ik), splat(0.5f), neg2(Sk));   // |X| - S
              const float2 dp = __ffma2_rn(__fmul2_rn(mp, ip), splat(0.5f), neg2(Sp));
              sc_num += dk.x * dk.x + dk.y * dk.y + dp.x * dp.x + dp.y * dp.y;
              sc_den += Sk.x * Sk.x + Sk.y * Sk.y + Sp.x * Sp.x + Sp.y * Sp.y;
            }
            const float2 PR = __fadd2_rn(YkR, YpR), PI = __fadd2_rn(YkI, neg2(YpI));
            const float2 DR = __fadd2_rn(YkR, neg2(YpR)), DI = __fadd2_rn(YkI, YpI);
            float2 vR = __ffma2_rn(GX, DR, PR);
            vR = __ffma2_rn(GY, DI, vR);
            float2 vI = __ffma2_rn(neg2(GX), DI, neg2(PI));
            vI = __ffma2_rn(GY, DR, vI);
            R[m] = vR; I[m] = vI;
            SR[m] = __ffma2_rn(PR, splat(2.0f), neg2(vR));
            SI[m] = __ffma2_rn(PI, splat(2.0f), vI);
          });
          if (l0) {   // k = 512 (self-paired): X = conj(Z[512]), conj(Z'2) = 2 Y
            const float S5 = spec_to_mag<SRC>(srow[512], g);
            const float2 X = make_float2(R[8].x, -I[8].x);
            const float m = X.x * X.x + X.y * X.y;
            const float im = rsqrt_fast(fmaxf(m, kTiny));
            const float fS = S5 * im;
            const float2 Y = make_float2(m > kTiny ? X.x * fS : S5, X.y * fS);
            if (SC) {
              const float d = m * im - S5;
              sc_num += d * d;
              sc_den += S5 * S5;
            }
            z512 = make_float2(2.0f * Y.x, 2.0f * Y.y);
          }
          float rr_[16], ri_[16];
          static_for<0, 8>([&](auto mc) {
            constexpr int m = decltype(mc)::value;
            rr_[2 * m] = __shfl_sync(0xffffffffu, SR[m].x, partner);
            rr_[2 * m + 1] = __shfl_sync(0xffffffffu, SR[m].y, partner);
            ri_[2 * m] = __shfl_sync(0xffffffffu, SI[m].x, partner);
            ri_[2 * m + 1] = __shfl_sync(0xffffffffu, SI[m].y, partner);
          });
          static_for<0, 8>([&](auto jc) {
            constexpr int j = decltype(jc)::value;
            const float ar = (j == 0) ? z512.x : rr_[(16 - 2 * j) & 15], ai = (j == 0) ? z512.y : ri_[(16 - 2 * j) & 15];
            R[8 + j] = make_float2(l0 ? ar : rr_[15 - 2 * j], l0 ? rr_[15 - 2 * j] : rr_[14 - 2 * j]);
            I[8 + j] = make_float2(l0 ? ai : ri_[15 - 2 * j], l0 ? ri_[15 - 2 * j] : ri_[14 - 2 * j]);
          });
        }

        // ------------------------------------------------------------------ 1024-point transform, 32 x 32
#pragma unroll 1
        for (int pass = 0; pass < 2; ++pass) {
          if (!(kProbe & 1)) fft32p(R, I);
          if (pass == 0) {
#pragma unroll
            for (int m = 0; m < 16 && !(kProbe & 2); ++m) {       // times W_1024^(lane * k2), k2 = 2m, 2m+1
              const float4 w = (kProbe & 8192) ? make_float4(0.7f, 0.6f, 0.5f + 1e-3f * m, 0.4f) : tw4[m * 32 + lane];
              const float2 WR = make_float2(w.x, w.y), WI = make_float2(w.z, w.w);
              const float2 nr = __ffma2_rn(R[m], WR, neg2(__fmul2_rn(I[m], WI)));
              I[m] = __ffma2_rn(R[m], WI, __fmul2_rn(I[m], WR));
              R[m] = nr;
            }
            __syncwarp();                                         // every lane is done with the buffer's previous contents
#pragma unroll
            for (int m = 0; m < 16 && !(kProbe & 4); ++m) {       // row k2: [re 0..31 | im 0..31], column = lane
              buf[(2 * m) * kRowFloats + lane] = R[m].x;
              buf[(2 * m) * kRowFloats + 32 + lane] = I[m].x;
              buf[(2 * m + 1) * kRowFloats + lane] = R[m].y;
              buf[(2 * m + 1) * kRowFloats + 32 + lane] = I[m].y;
            }
            __syncwarp();
#pragma unroll
            for (int jq = 0; jq < 8 && !(kProbe & 4); ++jq) {
              const float4 qr = *reinterpret_cast<const float4*>(&buf[lane * kRowFloats + 4 * jq]);
              const float4 qi = *reinterpret_cast<const float4*>(&buf[lane * kRowFloats + 32 + 4 * jq]);
              R[2 * jq] = make_float2(qr.x, qr.y); R[2 * jq + 1] = make_float2(qr.z, qr.w);
              I[2 * jq] = make_float2(qi.x, qi.y); I[2 * jq + 1] = make_float2(qi.z, qi.w);
            }
            __syncwarp();
            // the exchange buffer is idle until the next exchange: land this frame's |S| row (forward half) or the
            // next frame's input span (inverse half) in it, so that their latency hides behind the coming pass
            if ((kProbe & 128) || ((kProbe & 8) && half == 0) || ((kProbe & 512) && half == 0) || ((kProbe & 1024) && half == 1)) {
            } else if (half == 0) {
              s_off = span_to_smem_bulk_n<kF>(buf, spec_row0 + (long long)t * kF, a.spec, a.spec_end, lane, mbar_s, s_bulk);
            } else if (t + 1 < t_end) {
              x_off = span_issue(t + 1);
            }
          }
        }
      }  // halves

      // -------------------------------------------------------------------- window, overlap-add, emission: one pass
      // element n2 = conj(z'[lane + 32 n2]): sample 2q = Re, sample 2q+1 = -Im.  Pair q of this frame lives in ring slot
      // base + q.  Pairs q < count get their last term here and leave (* 1/(N wss) -> global, never stored back); pairs
      // q >= newq were not reached by the previous frame and are stored without a load; the rest is load-add-store.
      const bool last = t == T - 1;                                // the utterance's last frame flushes the whole window
      const int pn = ((t + 1) * HOP - WIN / 2) & 1;
      const int count = last ? G::kNP : (HOP + p - pn) >> 1;       // pairs [a0/2, a0(t+1)/2)
      const bool edge = last || t < G::kWarm || a0 < 0 || a0 + 2 * count > L;
      // pair q = lane + 32 n of this frame lives in ring slot (base + q) mod kRH: the rows from `wrap_row` on are addressed
      // from a second base pointer one ring length lower, so a row costs one compare and one select (immediate offsets)
      float2* const ring_a = ring + (base + lane);
      float2* const ring_b = ring_a - G::kRH;
      const int wrap_row = (G::kRH - base - lane + 31) >> 5;       // first row with base + lane + 32 n >= kRH
      // one row of 32 pairs; LOAD / EMIT: 0 = no lane, 1 = per lane (q < newq / q < count), 2 = every lane
      auto row = [&](auto nc, auto loadc, float2& v, float2*& sl) {
        constexpr int n = decltype(nc)::value, LOAD = decltype(loadc)::value;
        sl = (n >= wrap_row ? ring_b : ring_a) + 32 * n;
        v = make_float2(0.0f, 0.0f);
        if (LOAD == 2 || (LOAD == 1 && lane + 32 * n < newq)) v = *sl;
      };
      auto finish_row = [&](auto nc, auto emitc, auto zonec, float2 v, float2* sl, float tap_e, float tap_o) {
        constexpr int n = decltype(nc)::value, EMIT = decltype(emitc)::value;
        constexpr bool ZONE = decltype(zonec)::value;
        const float yr = (n & 1) ? R[n >> 1].y : R[n >> 1].x, yi = (n & 1) ? I[n >> 1].y : I[n >> 1].x;
        v.x = fmaf((kProbe & 32768) ? 0.5f + 1e-3f * n : tap_e, yr, v.x);
        v.y = fmaf((kProbe & 32768) ? -0.4f : -tap_o, yi, v.y);
        if (EMIT == 0 || (kProbe & 32)) {
          if (32 * n + 31 < G::kNP || lane + 32 * n < G::kNP) *sl = v;
        } else {
          const int i = a0 + 2 * lane + 64 * n;                    // sample index of v.x
          float2 sc2 = make_float2(pwx[2 * lane + 64 * n + 1 - p], pwx[2 * lane + 64 * n + 2 - p]);
          if (ZONE) {                                              // raw partial sums inside the run's head zone
            sc2.x = i < zone_end ? 1.0f : sc2.x;
            sc2.y = i + 1 < zone_end ? 1.0f : sc2.y;
          }
          if (EMIT == 2 || lane + 32 * n < count) *reinterpret_cast<float2*>(dst + i) = make_float2(v.x * sc2.x, v.y * sc2.y);
          else *sl = v;
        }
      };
      // rows in groups: all ring loads of a group first, then the arithmetic and the stores (loads and stores of the same
      // array cannot be reordered by the compiler; a load-modify-store chain per row would expose the latency 18 times)
      auto pass_rows = [&](auto zonec, auto edgec) {
        constexpr bool ZONE = decltype(zonec)::value, EDGE = decltype(edgec)::value;
        constexpr int kGroup = TTSA_RING_GROUP;                    // rows per group (a multiple of 4): 16-byte loads of the window taps
        static_for<0, (G::kRows + kGroup - 1) / kGroup>([&](auto gc) {
          constexpr int n0 = decltype(gc)::value * kGroup;
          float2 v[kGroup];
          float2* sl[kGroup];
          float tap_e[kGroup], tap_o[kGroup];
          static_for<0, kGroup / 4>([&](auto jc) {
            constexpr int j = decltype(jc)::value;
            if constexpr (n0 + 4 * j < G::kRows && !(kProbe & 32768)) {
              const float4 e4 = *reinterpret_cast<const float4*>(pe + n0 + 4 * j), o4 = *reinterpret_cast<const float4*>(po + n0 + 4 * j);
              tap_e[4 * j] = e4.x; tap_e[4 * j + 1] = e4.y; tap_e[4 * j + 2] = e4.z; tap_e[4 * j + 3] = e4.w;
              tap_o[4 * j] = o4.x; tap_o[4 * j + 1] = o4.y; tap_o[4 * j + 2] = o4.z; tap_o[4 * j + 3] = o4.w;
            } else {
              tap_e[4 * j] = tap_e[4 * j + 1] = tap_e[4 * j + 2] = tap_e[4 * j + 3] = 0.0f;
              tap_o[4 * j] = tap_o[4 * j + 1] = tap_o[4 * j + 2] = tap_o[4 * j + 3] = 0.0f;
            }
          });
          static_for<0, kGroup>([&](auto jc) {
            constexpr int n = n0 + decltype(jc)::value;
            if constexpr (n < G::kRows) {
              // which lanes of this row still hold partial sums of earlier frames
              constexpr int LOAD = (ZONE || EDGE) ? 1 : (32 * n + 31 < G::kNP - G::kCntMax ? 2 : (32 * n < G::kNP - G::kCntMin ? 1 : 0));
              row(IntC<n>{}, IntC<LOAD>{}, v[decltype(jc)::value], sl[decltype(jc)::value]);
            }
          });
          static_for<0, kGroup>([&](auto jc) {
            constexpr int n = n0 + decltype(jc)::value;
            if constexpr (n < G::kRows) {
              constexpr int EMIT = EDGE ? 0 : (32 * n + 31 < G::kCntMin ? 2 : (32 * n < G::kCntMax ? 1 : 0));
              finish_row(IntC<n>{}, IntC<EMIT>{}, zonec, v[decltype(jc)::value], sl[decltype(jc)::value], tap_e[decltype(jc)::value], tap_o[decltype(jc)::value]);
            }
          });
        });
      };
      if (TTSA_ZONE_DEFER && zone_state == 2) { publish_flag(a.wps_flags + wi, epoch, l0); zone_state = 0; }
      if (kProbe & 16) {
      } else if (!edge) {
        if (a0 >= zone_end && newq > 0) pass_rows(IntC<0>{}, IntC<0>{});   // steady state
        else pass_rows(IntC<1>{}, IntC<0>{});                              // head zone of a run / its first frame
      } else {
        // utterance edges (first kWarm frames, last frame): everything goes through the ring, then a slow emission with
        // the window sum over the frames that exist
        pass_rows(IntC<0>{}, IntC<1>{});
        __syncwarp();
#pragma unroll 1
        for (int e = lane; e < count && !(kProbe & 32); e += 32) {
          int sl = base + e; sl = sl >= G::kRH ? sl - G::kRH : sl;
          const float2 v = ring[sl];
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int i = a0 + 2 * e + h;
            const float val = h ? v.y : v.x;
            if (i >= 0 && i < L) dst[i] = i < zone_end ? val : val * edge_inv(i, T, L);   // raw partial sum inside the head zone
          }
        }
      }
      __syncwarp();
      newq = G::kNP - count;
      base += count; base = base >= G::kRH ? base - G::kRH : base;
      // the head zone is stored: the owner of the frames before the cut may finish it.  The flag is published one frame later
      // (before that frame's stores): the fence of the release then finds this frame's stores long complete instead of
      // waiting for them (2.6 % of all warp time sat in that fence)
      if (zone_state == 1 && a0 + 2 * count >= zone_end) {
        zone_state = 2;
        if (!TTSA_ZONE_DEFER) { publish_flag(a.wps_flags + wi, epoch, l0); zone_state = 0; }
      }
    }  // frames of the run
    if (zone_state == 2) { publish_flag(a.wps_flags + wi, epoch, l0); zone_state = 0; }
    WPS_STAMP(2);

    if constexpr (SC) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        sc_num += __shfl_xor_sync(0xffffffffu, sc_num, o);
        sc_den += __shfl_xor_sync(0xffffffffu, sc_den, o);
      }
      if (lane == 0) {
        float* const acc = a.sc_acc + 2 * ((size_t)it * bd.B + u);     // [iteration][utterance][num, den]
        atomicAdd(acc, sc_num);
        atomicAdd(acc + 1, sc_den);
      }
    }
    // A run that ends inside the utterance holds the partial sums of the next run's head zone in its ring: wait until
    // that run's warp (the next one with a non-empty range; possibly in the next CTA) has stored its share -- it does so
    // within its first kWarm + 1 frames, and waits for nobody before -- then finish the zone:
    // out = (its raw partial + this ring) / (N wss).
    if (t_end < T) {
      const int s0z = t_end * HOP - WIN / 2;
      const int a0z = s0z - (s0z & 1);
      const int zend = (t_end - 1) * HOP + WIN / 2;
      constexpr int kZoneIters = ((G::kZone + 3) / 2 + 31) / 32;
      int nx = wi + 1;
      while (wp.cut[nx + 1] == wp.cut[nx]) ++nx;                   // the cut list ends at the total frame count > cut[nx]
      wait_flag_ge(a.wps_flags + nx, epoch, l0);
      WPS_STAMP(3);
      const bool all_frames = t_end >= G::kWarm && t_end + G::kWarm < T;
      float2 v[kZoneIters], o[kZoneIters];
#pragma unroll
      for (int k = 0; k < kZoneIters; ++k) {
        const int e = lane + 32 * k;
        const int i = a0z + 2 * e;
        int sl = base + e; sl = sl >= G::kRH ? sl - G::kRH : sl;
        v[k] = make_float2(0.0f, 0.0f); o[k] = v[k];
        if (i < zend && i >= 0 && i < L) {
          v[k] = ring[sl];
          o[k] = ld_volatile_f2(dst + i);
        }
      }
#pragma unroll
      for (int k = 0; k < kZoneIters; ++k) {
        const int e = lane + 32 * k;
        const int i = a0z + 2 * e;
        if (i < zend && i >= 0 && i < L) {
          float2 r2 = o[k];
          if (all_frames) {
            int r = (i - s0z) % HOP; r = r < 0 ? r + HOP : r;
            r2.x = (o[k].x + v[k].x) * pwx[r + 1];
            if (i + 1 < zend) r2.y = (o[k].y + v[k].y) * pwx[r + 2];
          } else {
            r2.x = (o[k].x + v[k].x) * edge_inv(i, T, L);
            if (i + 1 < zend && i + 1 < L) r2.y = (o[k].y + v[k].y) * edge_inv(i + 1, T, L);
          }
          *reinterpret_cast<float2*>(dst + i) = r2;
        }
      }
    }
  }  // runs
  WPS_STAMP(4);
  if (n_it > 1 && fa < fb) {                                       // this warp's part of iteration `it` is complete and visible
    publish_flag(a.wps_done + wi, epoch, l0);
  }
  WPS_STAMP(5);
  }  // iterations
}

}  // namespace ttsa
