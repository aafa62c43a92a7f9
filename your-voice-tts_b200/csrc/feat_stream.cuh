// Feature extraction, "warp stream" form (utils/audio.py:138-152, datasets/TTSDataset.py:191-192): spectrogram and
// melspectrogram of a batch from ONE STFT, every warp walking a private run of consecutive frames with no CTA-wide
// barrier -- the analysis half of gl_stream.cuh (same span copies, window taps per lane, packed 32 x 32 transform), which
// needs neither an overlap-add ring nor a hand-over between runs:
//
//   per frame (one warp):  bulk-copied input span -> pre-emphasis, window -> FFT (32 x 32, one exchange)
//                          -> real-FFT post-processing on conjugate pairs -> |X|
//                          -> dB, normalise -> linear row straight from registers (coalesced 128-byte stores)
//                          -> |X| row in shared memory -> banded Slaney mel contraction -> dB, normalise -> mel row
//
// The next frame's span lands in the upper half of the warp's exchange buffer (above the 1025 magnitudes) right after the
// exchange, so its latency hides behind the second transform pass and the output stage.
//
// Pre-emphasis (utils/audio.py:128-131) is applied to the raw signal BEFORE the reflect padding of librosa.stft:
// z[j] = y[j] - c y[j-1] (z[0] = y[0]), frames are cut from reflect-padded z.  For a padded index i the neighbour that
// plays y[j-1] is the padded sample at i - 1 inside the signal and at i + 1 in the mirrored parts, so the kernel works on
// the reflect-padded RAW span and picks the neighbour per sample (interior frames: always i - 1).
#pragma once
#include "gl_stream.cuh"

namespace ttsa {

constexpr int kFeatSpanAt = 1040;    // floats: where a span lands in the warp's buffer (the magnitudes use [0, 1025))

template <int HOP, int WIN>
struct FeatGeo {
  using G = WpsGeo<HOP, WIN>;
  static constexpr int kSpan = WIN + 6;                          // samples [a0 - 2, a0 + WIN + 4): one neighbour on either side
  static constexpr int sm_img = kWpsWarps * kBufFloats;          // the warp-stream table image (Tables::wps_image): tw4 | g4 | wA | wB | pwx
  static constexpr int sm_tw = sm_img;
  static constexpr int sm_g = sm_tw + 2048;
  static constexpr int sm_wA = sm_g + 1024;
  static constexpr int sm_wB = sm_wA + 32 * G::kWS;
  static constexpr int sm_mbar = sm_img + G::image_floats;       // table barrier, then one barrier per warp (8 bytes each)
  static constexpr int sm_mel = sm_mbar + 4 + 2 * kWpsWarps;     // compact mel basis (Geo::mel_smem_floats floats; 0: read from global)
  // the span must fit above the magnitude row (300 / 1200 does not: the tile kernel serves it)
  static constexpr bool kFits = kFeatSpanAt + 12 + kSpan + 3 <= kBufFloats && (sm_mel + 2048) * 4 <= 227 * 1024;
};

// LIN / MEL: which outputs the launch produces (compile-time, so that the output stage has no branch the compiler must
// treat as possibly divergent around the warp shuffles)
template <int HOP, int WIN, bool LIN, bool MEL>
__global__ void __launch_bounds__(kWpsThreads, 1)
feat_stream_kernel(const Geo g, const Tables tb, const BatchDev bd, const FrameArgs a, const int total_frames) {
  using G = WpsGeo<HOP, WIN>;
  using FG = FeatGeo<HOP, WIN>;
  extern __shared__ __align__(16) float smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = warp_index();
  float* const buf = smem + warp * kBufFloats;
  const float4* const tw4 = reinterpret_cast<const float4*>(smem + FG::sm_tw);
  const float4* const g4 = reinterpret_cast<const float4*>(smem + FG::sm_g);
  const float* const wA = smem + FG::sm_wA;
  const float* const wB = smem + FG::sm_wB;

  // ---- prologue: table image by one bulk copy, the compact mel basis by ordinary loads
  const unsigned mbar = (unsigned)__cvta_generic_to_shared(smem + FG::sm_mbar);
  if (tid == 0) {
    const unsigned dst = (unsigned)__cvta_generic_to_shared(smem + FG::sm_img);
    constexpr unsigned bytes = (unsigned)G::image_floats * 4u;
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar) : "memory");
    for (int i = 0; i < kWpsWarps; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar + 16 + 8 * i) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(tb.wps_image), "r"(bytes), "r"(mbar) : "memory");
  }
  if (MEL) {
    if (g.mel_seg_pairs[0] > 0) {
      const int nw_ = 160 * (g.mel_seg_pairs[0] + g.mel_seg_pairs[1] + g.mel_seg_pairs[2]) + 96;
      unsigned* const d = reinterpret_cast<unsigned*>(smem + FG::sm_mel);
      for (int i = tid; i < nw_; i += kWpsThreads) d[i] = tb.mel_seg[i];
      if (lane == 0) smem[FG::sm_mel + nw_ + warp * kMelSegScratch + 192] = 0.0f;      // the zero cell of this warp's scratch row
    } else if (g.mel_steps > 0) for (int i = tid; i < g.mel_steps * 128 + 96; i += kWpsThreads) smem[FG::sm_mel + i] = tb.mel_sched[i];
    else for (int i = tid; i < g.mel_smem_floats; i += kWpsThreads) smem[FG::sm_mel + i] = tb.mel_compact[i];
  }
  __syncthreads();
  mbar_wait(mbar, 0);

  const int nw = gridDim.x * kWpsWarps, wi = blockIdx.x * kWpsWarps + warp;
  const int fa = (int)((long long)wi * total_frames / nw), fb = (int)((long long)(wi + 1) * total_frames / nw);
  if (fa >= fb) return;                                            // (no CTA-wide barrier below)
  int u = 0;
  { int lo = 0, hi = bd.B; while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (bd.tsum[mid] <= fa) lo = mid; else hi = mid; } u = lo; }

  const unsigned mbar_x = mbar + 16 + 8 * warp;
  unsigned ph_x = 0;
  const int partner = (32 - lane) & 31;
  const bool l0 = lane == 0;
  const float pc = a.preemph ? g.preemph : 0.0f;

  int f = fa;
  while (f < fb) {
    while (bd.tsum[u + 1] <= f) ++u;                               // skips empty utterances
    const int tsu = uni(bd.tsum[u]);
    const int T = uni(bd.T[u]);
    const int t_begin = f - tsu;
    const int t_end = min(T, fb - tsu);
    f = tsu + t_end;
    const int L = uni(bd.wav_len[u]);
    if (L <= 0) continue;
    const float* __restrict__ src = a.wav_in + uni(bd.wav_off[u]);
    const long long row0 = uni(bd.frame_off[u]);

    // span of frame t: raw samples [a0 - 2, a0 + WIN + 4), a0 = the frame's first sample made even.  0: inside the utterance
    // (one bulk copy), 1: first / last frames of a long utterance (the part that exists by a bulk copy, the mirrored part
    // filled in from shared memory), 2: short utterances (index-mapped loads; pre-emphasis applied while loading)
    const bool edge_async = L >= 2 * FG::kSpan;
    auto frame_a0 = [&](int t) { const int s0 = t * HOP - WIN / 2; return s0 - (s0 & 1); };
    auto span_mode = [&](int t) {
      const int b0 = frame_a0(t) - 2;
      return (b0 >= 0 && b0 + FG::kSpan <= L) ? 0 : (edge_async ? 1 : 2);
    };
    auto span_issue = [&](int t) {                                 // returns the position of sample a0 in buf
      const int a0 = frame_a0(t), b0 = a0 - 2;
      float* const land = buf + kFeatSpanAt;
      const int mode = span_mode(t);
      if (mode == 0) return kFeatSpanAt + 2 + span_to_smem_bulk_inside<FG::kSpan>(land, src + b0, lane, mbar_x);
      if (mode == 1) {
        const int A = (b0 - 8) & ~3;                               // sample that lands at land[0]
        if (lane == 0) {
          const int c_lo = max(A, 0), c_hi = (min(b0 + FG::kSpan, L) + 3) & ~3;
          const unsigned bytes = (unsigned)(c_hi - c_lo) << 2;
          const unsigned d = (unsigned)__cvta_generic_to_shared(land + (c_lo - A));
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar_x), "r"(bytes) : "memory");
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                       ::"r"(d), "l"(src + c_lo), "r"(bytes), "r"(mbar_x) : "memory");
        }
        return kFeatSpanAt + (a0 - A);
      }
      // np.pad(..., mode='reflect') as an index map, any number of folds; z = y[r] - c y[r - 1] goes into the buffer
#pragma unroll 1
      for (int m = lane; m < FG::kSpan; m += 32) {
        const int r = reflect_index(b0 + m, L);
        float v = __ldg(src + r);
        if (r > 0) v = fmaf(-pc, __ldg(src + r - 1), v);
        land[m] = v;
      }
      return kFeatSpanAt + 2;
    };
    // mode 1: fill the samples outside [0, L) from their mirror images (one fold; a mirror image outside the copy belongs to
    // an element no tap uses)
    auto fill_reflected = [&](int a0, int x_off) {
      const int b0 = a0 - 2, A = a0 - (x_off - kFeatSpanAt);
      float* const land = buf + kFeatSpanAt;
      const int c_lo = max(A, 0), c_hi = min(b0 + FG::kSpan, L);
#pragma unroll 1
      for (int idx = b0 + lane; idx < 0; idx += 32) land[idx - A] = (-idx < c_hi) ? land[-idx - A] : 0.0f;
#pragma unroll 1
      for (int idx = L + lane; idx < b0 + FG::kSpan; idx += 32) {
        const int r = 2 * (L - 1) - idx;
        land[idx - A] = (r >= c_lo) ? land[r - A] : 0.0f;
      }
      __syncwarp();
    };

    int x_off = span_issue(t_begin);
#pragma unroll 1
    for (int t = t_begin; t < t_end; ++t) {
      const int s0 = t * HOP - WIN / 2;
      const int p = s0 & 1;
      const int a0 = s0 - p;
      const int mode = span_mode(t);
      // window of the even / odd sample of pair q = lane + 32 n: pe[n], po[n] (an odd frame starts one sample early)
      const float* const pe = (p ? wB : wA) + lane * G::kWS;
      const float* const po = (p ? wA : wB + G::kWS) + lane * G::kWS;

      float2 R[16], I[16];
      // ------------------------------------------------------------------ input span -> pre-emphasis -> windowed packed frame
      if (mode != 2) { mbar_wait(mbar_x, ph_x); ph_x ^= 1u; }
      __syncwarp();
      if (mode == 1) fill_reflected(a0, x_off);
      {
        const float* const xs = buf + x_off;                       // xs[j] = padded raw sample a0 + j (mode 2: pre-emphasised already)
#pragma unroll
        for (int m = 0; m < 16; ++m) {
          if (m < G::kMH) {
            float2 za = make_float2(0.0f, 0.0f), zb = make_float2(0.0f, 0.0f);
            static_for<0, 2>([&](auto hc) {
              constexpr int h = decltype(hc)::value;
              const int q = lane + 64 * m + 32 * h;
              if (2 * m + h < G::kRows) {
                const float2 x = *reinterpret_cast<const float2*>(xs + 2 * q);
                float2 z = x;
                if (mode == 0) {                                   // inside the signal: the neighbour is the sample before
                  z.x = fmaf(-pc, xs[2 * q - 1], x.x);
                  z.y = fmaf(-pc, x.x, x.y);
                } else if (mode == 1) {
                  const int i0 = a0 + 2 * q, i1 = i0 + 1;
                  const float n0 = (i0 <= 0 || i0 >= L) ? (i0 == 0 ? 0.0f : x.y) : xs[2 * q - 1];
                  const float n1 = (i1 <= 0 || i1 >= L) ? (i1 == 0 ? 0.0f : xs[2 * q + 2]) : x.x;
                  z.x = fmaf(-pc, n0, x.x);
                  z.y = fmaf(-pc, n1, x.y);
                }
                // pairs past the window hold other data: force zeros (their taps are zero, 0 * NaN must not leak)
                if (64 * m + 32 * h + 31 >= WIN / 2 && q >= WIN / 2 + p) z = make_float2(0.0f, 0.0f);
                if (h == 0) za = z; else zb = z;
              }
            });
            const float2 we = *reinterpret_cast<const float2*>(pe + 2 * m), wo = *reinterpret_cast<const float2*>(po + 2 * m);
            R[m] = __fmul2_rn(make_float2(za.x, zb.x), we);
            I[m] = __fmul2_rn(make_float2(za.y, zb.y), wo);
          } else {
            R[m] = make_float2(0.0f, 0.0f);
            I[m] = make_float2(0.0f, 0.0f);
          }
        }
      }
      __syncwarp();                                                // the landing zone becomes the exchange buffer

      // -------------------------------------------------------------------- 1024-point transform, 32 x 32
#pragma unroll 1
      for (int pass = 0; pass < 2; ++pass) {
        fft32p(R, I);
        if (pass == 0) {
#pragma unroll
          for (int m = 0; m < 16; ++m) {                           // times W_1024^(lane * k2), k2 = 2m, 2m+1
            const float4 w = tw4[m * 32 + lane];
            const float2 WR = make_float2(w.x, w.y), WI = make_float2(w.z, w.w);
            const float2 nr = __ffma2_rn(R[m], WR, neg2(__fmul2_rn(I[m], WI)));
            I[m] = __ffma2_rn(R[m], WI, __fmul2_rn(I[m], WR));
            R[m] = nr;
          }
#pragma unroll
          for (int m = 0; m < 16; ++m) {                           // row k2: [re 0..31 | im 0..31], column = lane
            buf[(2 * m) * kRowFloats + lane] = R[m].x;
            buf[(2 * m) * kRowFloats + 32 + lane] = I[m].x;
            buf[(2 * m + 1) * kRowFloats + lane] = R[m].y;
            buf[(2 * m + 1) * kRowFloats + 32 + lane] = I[m].y;
          }
          __syncwarp();
#pragma unroll
          for (int jq = 0; jq < 8; ++jq) {
            const float4 qr = *reinterpret_cast<const float4*>(&buf[lane * kRowFloats + 4 * jq]);
            const float4 qi = *reinterpret_cast<const float4*>(&buf[lane * kRowFloats + 32 + 4 * jq]);
            R[2 * jq] = make_float2(qr.x, qr.y); R[2 * jq + 1] = make_float2(qr.z, qr.w);
            I[2 * jq] = make_float2(qi.x, qi.y); I[2 * jq + 1] = make_float2(qi.z, qi.w);
          }
          __syncwarp();
          // the buffer is idle until the next frame's exchange except for the magnitude row [0, 1025): land the next span above it
          if (t + 1 < t_end) x_off = span_issue(t + 1);
        }
      }

      // ---------------------------------------------------------------------- spectrum -> |X| -> outputs
      // X[k] = (E2 + G_k D2) / 2, X[1024 - k] = conj(E2 - G_k D2) / 2 on conjugate pairs (k, 1024 - k) of the packed transform
      const long long row = row0 + t;
      float* const lout = LIN ? a.lin_out + row * kF : nullptr;
      constexpr bool want_mel = MEL;
      float* const magbuf = buf;
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
        const float2 XpR = __ffma2_rn(E2R, splat(2.0f), neg2(XkR));
        float2 XpI = __ffma2_rn(E2I, splat(-2.0f), XkI);
        if (m == 0 && l0) { XkI.x = 0.0f; XpI.x = 0.0f; }            // DC and Nyquist are real
        // |X| = |2 X| / 2
        const float2 mk = __fmul2_rn(__ffma2_rn(XkI, XkI, __fmul2_rn(XkR, XkR)), splat(0.25f));
        const float2 mp = __fmul2_rn(__ffma2_rn(XpI, XpI, __fmul2_rn(XpR, XpR)), splat(0.25f));
        const float a0m = sqrt_fast(mk.x), a1m = sqrt_fast(mk.y), a2m = sqrt_fast(mp.x), a3m = sqrt_fast(mp.y);
        if (LIN) {
          lout[k0] = amp_to_norm_db(a0m, g);
          lout[k0 + 32] = amp_to_norm_db(a1m, g);
          lout[1024 - k0] = amp_to_norm_db(a2m, g);
          lout[992 - k0] = amp_to_norm_db(a3m, g);
        }
        if (want_mel) {
          magbuf[k0] = a0m; magbuf[k0 + 32] = a1m; magbuf[1024 - k0] = a2m; magbuf[992 - k0] = a3m;
        }
      });
      {   // k = 512: X = conj(Z[512]), lane 0's.  Every lane computes, lane 0 stores: a divergent branch here is not
          // reconverged before the end of the frame and the mel contraction below would run once per half of the warp
        const float am = sqrt_fast(R[8].x * R[8].x + I[8].x * I[8].x);
        const float dbv = amp_to_norm_db(am, g);
        if (LIN && l0) lout[512] = dbv;
        const float am0 = __shfl_sync(0xffffffffu, am, 0);
        if (want_mel) magbuf[512] = am0;                          // every lane, same value: no lane-dependent branch
      }
      if (want_mel) {
        __syncwarp();
        float* const out = a.mel_out + row * g.num_mels;
        if (g.mel_seg_pairs[0] > 0) {
          // segment schedule (host_tables.hpp, mel_segment_schedule): every bin is read once and feeds the falling tap of one
          // filter and the rising tap of the next; three slots of cells, then <= 4 partial sums per filter through the scratch row
          const int np_ = g.mel_seg_pairs[0] + g.mel_seg_pairs[1] + g.mel_seg_pairs[2];
          const float4* const wq = reinterpret_cast<const float4*>(smem + FG::sm_mel) + lane;
          const unsigned* const iq = reinterpret_cast<const unsigned*>(smem + FG::sm_mel + 128 * np_) + lane;
          const unsigned* const comb = iq + 32 * np_;
          float* const part = smem + FG::sm_mel + 160 * np_ + 96 + warp * kMelSegScratch;
          const char* const mb = reinterpret_cast<const char*>(magbuf);
          int pr = 0;
#pragma unroll
          for (int sl = 0; sl < 3; ++sl) {
            float accd = 0.0f, accu = 0.0f;
            const int pend = pr + g.mel_seg_pairs[sl];
#pragma unroll 4
            for (; pr < pend; ++pr) {
              const float4 w = wq[32 * pr];
              const unsigned ix = iq[32 * pr];
              const float ma = *reinterpret_cast<const float*>(mb + (ix & 0xffffu));
              const float mg = *reinterpret_cast<const float*>(mb + (ix >> 16));
              accd = fmaf(w.x, ma, accd); accu = fmaf(w.y, ma, accu);
              accd = fmaf(w.z, mg, accd); accu = fmaf(w.w, mg, accu);
            }
            *reinterpret_cast<float2*>(part + 2 * (32 * sl + lane)) = make_float2(accd, accu);
          }
          __syncwarp();
#pragma unroll
          for (int r = 0; r < 3; ++r) {
            const unsigned cw = comb[32 * r];
            const float sum = ((part[cw & 255u] + part[(cw >> 8) & 255u]) + part[(cw >> 16) & 255u]) + part[cw >> 24];
            const float v = amp_to_norm_db(sum, g);
            if (32 * r + lane < g.num_mels) out[32 * r + lane] = v;
          }
        } else if (g.mel_steps > 0) {
          // lane schedule (built by the host with the plan): up to three filters per lane, about sum(taps) / 32 steps, the 32
          // magnitude reads of a step in 32 different banks; one 16-byte load, one magnitude load, three FMAs per step
          const float4* const sch = reinterpret_cast<const float4*>(smem + FG::sm_mel) + lane;
          float acc0 = 0.0f, acc1 = 0.0f, acc2 = 0.0f;
#pragma unroll 8
          for (int st = 0; st < g.mel_steps; ++st) {
            const float4 e = sch[32 * st];
            const float mg = magbuf[__float_as_int(e.w)];
            acc0 = fmaf(e.x, mg, acc0);
            acc1 = fmaf(e.y, mg, acc1);
            acc2 = fmaf(e.z, mg, acc2);
          }
          const int* const fid = reinterpret_cast<const int*>(smem + FG::sm_mel + g.mel_steps * 128) + lane;
          const int f0 = fid[0], f1 = fid[32], f2 = fid[64];
          const float v0 = amp_to_norm_db(acc0, g), v1 = amp_to_norm_db(acc1, g), v2 = amp_to_norm_db(acc2, g);
          if (f0 >= 0) out[f0] = v0;
          if (f1 >= 0) out[f1] = v1;
          if (f2 >= 0) out[f2] = v2;
        } else if (g.mel_smem_floats > 0) {
          // compact basis in shared memory: (first tap, first bin) per filter, one extra entry, then the taps
          const int2* const mdesc = reinterpret_cast<const int2*>(smem + FG::sm_mel);
          const float* const mval = smem + FG::sm_mel + 2 * (g.num_mels + 1);
          for (int mm = lane; mm < g.num_mels; mm += 32) {
            const int2 d0 = mdesc[mm];
            const int cnt = mdesc[mm + 1].x - d0.x;
            const float* mv = mval + d0.x;
            const float* mg = magbuf + d0.y;
            float acc = 0.0f;
            for (int c = 0; c < cnt; ++c) acc = fmaf(mv[c], mg[c], acc);
            out[mm] = amp_to_norm_db(acc, g);
          }
        } else {
          for (int mm = lane; mm < g.num_mels; mm += 32) {
            const int lo = tb.mel_lo[mm], cnt = tb.mel_cnt[mm];
            const float* mv = tb.mel_val + mm * tb.mel_ld;
            float acc = 0.0f;
            for (int c = 0; c < cnt; ++c) acc = fmaf(__ldg(mv + c), magbuf[lo + c], acc);
            out[mm] = amp_to_norm_db(acc, g);
          }
        }
      }
      __syncwarp();                                                // the magnitude row is consumed before the next exchange
    }  // frames of the run
  }  // runs
}

}  // namespace ttsa
