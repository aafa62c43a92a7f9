// Griffin-Lim iteration as a barrier-free warp chain (standard geometry class).
//
//   y -> [window, FFT] -> |S| e^{j angle X} -> [iFFT, window, overlap-add, / wss] -> y'      utils/audio.py:186-188
//
// One warp owns one frame at a time and never meets a CTA-wide barrier inside a segment.  A CTA works through a run of
// consecutive frames of one utterance; frame number s of the run belongs to warp s % 8.  Each warp loads its frame's
// samples straight from global memory (L2 resident: every sample is read by ~4 frames), transforms, projects,
// transforms back, and then adds the windowed result into a circular overlap-add accumulator in shared memory.  The
// adds happen in STRICT FRAME ORDER: warp w waits on a named barrier for the warp that owns the previous frame, adds,
// and releases the next warp.  The order makes the sum deterministic, and it means that right after frame t has been
// added the `hop` samples [t*hop - off0, (t+1)*hop - off0) are final: the same warp scales them by 1/(N * wss), stores
// them to HBM and clears them.  In steady state the warps run skewed by one add each and nobody waits.
//
// A segment that does not start at frame 0 first recomputes the (win-1)/hop preceding frames (no output) so that its
// first samples receive every overlapping contribution; each output sample is written exactly once, by one CTA.
#pragma once
#include "frame_core.cuh"

namespace ttsa {

constexpr int kAccLen = 4096;                 // circular accumulator, samples; needs >= win + 7*hop (checked on host)
constexpr int kAccPlane = kAccLen / 2 + 16;   // even-sample plane, then odd-sample plane 16 banks further

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void named_bar_arrive(int id, int nthreads) {
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

struct ChainSmem {      // float offsets
  int acc, wE, wO, pw, tw, g, total;
};

template <int SRC, bool SC>
__global__ void __launch_bounds__(kThreads, 2)
gl_chain_kernel(const Geo g, const Tables tb, const BatchDev bd, const FrameArgs a, const ChainSmem sm, const long long total_frames) {
  extern __shared__ __align__(16) float smem[];
  constexpr int NZ = 20;
  constexpr int ND = 5;
  constexpr float kInvN = 1.0f / (float)kNfft;
  constexpr int kMask = kAccLen / 2 - 1;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  float* const buf = smem + warp * kBufFloats;
  float* const accE = smem + sm.acc;
  float* const accO = accE + kAccPlane;
  const float2* const wE2 = reinterpret_cast<const float2*>(smem + sm.wE);
  const float2* const wO2 = reinterpret_cast<const float2*>(smem + sm.wO);
  float* const pw = smem + sm.pw;
  const float4* const tw4 = reinterpret_cast<const float4*>(smem + sm.tw);
  const float4* const g4 = reinterpret_cast<const float4*>(smem + sm.g);

  for (int i = tid; i < 512; i += kThreads) reinterpret_cast<float4*>(smem + sm.tw)[i] = tb.tw4[i];
  for (int i = tid; i < 256; i += kThreads) reinterpret_cast<float4*>(smem + sm.g)[i] = tb.g4[i];
  for (int i = tid; i < NZ * 16; i += kThreads) {
    reinterpret_cast<float2*>(smem + sm.wE)[i] = tb.wE2[i];
    reinterpret_cast<float2*>(smem + sm.wO)[i] = tb.wO2[i];
  }
  for (int i = tid; i < g.hop; i += kThreads) pw[i] = tb.pw[i] * kInvN;     // 1/wss and the 1/n_fft of the inverse FFT

  // contiguous range of this CTA over the flattened (utterance, frame) list
  const long long f_lo = (long long)blockIdx.x * total_frames / gridDim.x;
  const long long f_hi = (long long)(blockIdx.x + 1) * total_frames / gridDim.x;
  int u = 0;
  {
    int lo = 0, hi = bd.B;   // largest u with frame_off[u] <= f_lo
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (bd.frame_off[mid] <= f_lo) lo = mid; else hi = mid;
    }
    u = lo;
  }

  long long fpos = f_lo;
  while (fpos < f_hi) {
    while (u + 1 < bd.B && fpos >= bd.frame_off[u + 1]) ++u;
    const int T = bd.T[u];
    const long long frow0 = bd.frame_off[u];
    const int ta = (int)(fpos - frow0);
    const long long seg_end = f_hi < bd.frame_off[u + 1] ? f_hi : bd.frame_off[u + 1];
    const int tb_ = (int)(seg_end - frow0);
    fpos = seg_end;
    const int L = bd.wav_len[u];                 // hop*(T-1)
    if (L <= 0) continue;
    const long long woff = bd.wav_off[u];
    const float* __restrict__ src = a.wav_in + woff;
    float* __restrict__ dst = a.wav_out + woff;
    const int n_warm = ta < g.nwarm ? ta : g.nwarm;
    const int t_first = ta - n_warm;
    const int n_seq = tb_ - t_first;
    float sc_num = 0.0f, sc_den = 0.0f;

    __syncthreads();                               // previous segment completely drained
    for (int i = tid; i < 2 * kAccPlane; i += kThreads) accE[i] = 0.0f;
    __syncthreads();

    // Asynchronous staging of a frame's win samples into the warp's (idle) exchange buffer.  Interior frames use
    // 16-byte chunks; frames that touch the reflect-padded edges copy element-wise through the reflect index map.
    auto stage_frame = [&](int t) -> int {
      const int i_start = t * g.hop - g.off0;
      if (i_start >= 0 && i_start + g.win <= L)
        return span_to_smem_async(buf, src + i_start, g.win, a.wav_in, a.wav_end, lane);
      for (int m = lane; m < g.win; m += 32) cp_async4(buf + m, src + reflect_index(i_start + m, L));
      return 0;
    };
    int x_off = 0;
    if (warp < n_seq) x_off = stage_frame(t_first + warp);

    for (int s = warp; s < n_seq; s += kNF) {
      const int t = t_first + s;
      const bool own = t >= ta;                    // warm-up frames only feed the accumulator
      const long long row = frow0 + t;
      const int a0 = t * g.hop;                    // accumulator coordinate of tap 0 (= sample index + off0)
      const int i_start = a0 - g.off0;
      float2 R[16], I[16];
      int s_off = 0;

      // ---- window the staged frame: element n2 of this lane = x[2q] + j x[2q+1], q = lane + 32 n2
      cp_async_wait_all();
      __syncwarp();
      {
        const float* xs = buf + x_off + 2 * lane;
#pragma unroll
        for (int m = 0; m < 16; ++m) {
          if (2 * m < NZ) {
            const int q0 = lane + 64 * m, q1 = q0 + 32;
            const float x0 = 2 * q0 < g.win ? xs[128 * m] : 0.0f, y0 = 2 * q0 + 1 < g.win ? xs[128 * m + 1] : 0.0f;
            const float x1 = 2 * q1 < g.win ? xs[128 * m + 64] : 0.0f, y1 = 2 * q1 + 1 < g.win ? xs[128 * m + 65] : 0.0f;
            R[m] = __fmul2_rn(make_float2(x0, x1), wE2[m * 32 + lane]);
            I[m] = __fmul2_rn(make_float2(y0, y1), wO2[m * 32 + lane]);
          } else {
            R[m] = make_float2(0.0f, 0.0f);
            I[m] = make_float2(0.0f, 0.0f);
          }
        }
      }

      // ---- analysis transform, projection, synthesis transform: ONE rolled copy of the transform body
      //      (ifft(Z) = conj(fft(conj(Z))): the per-bin step emits conj(Z'), the add below takes -Im)
#pragma unroll 1
      for (int half = 0; half < 2; ++half) {
        if (half == 1) gl_bin_step<SRC, SC>(R, I, buf + s_off, g4, g, lane, own, sc_num, sc_den);
#pragma unroll 1
        for (int pass = 0; pass < 2; ++pass) {
          transform_pass(R, I, buf, tw4, lane, pass == 0);
          if (pass == 0) {
            // the exchange buffer is idle until the next transform: start the copy that is needed next -- this
            // frame's |S| row, or this warp's next frame -- so that its latency hides behind the second pass
            if (half == 0) s_off = span_to_smem_async(buf, a.spec + row * kF, kF, a.spec, a.spec_end, lane);
            else if (s + kNF < n_seq) x_off = stage_frame(t + kNF);
          }
        }
      }

      // ---- ordered overlap-add: element n2 = conj(z'[lane + 32 n2]); sample 2q = Re, sample 2q+1 = -Im
      if (s > 0) named_bar_sync(1 + ((warp + kNF - 1) & (kNF - 1)), 64);      // frame s-1 has been added
      {
        // the critical section of the chain: all loads, then all FMAs, then all stores (no false dependences)
        const int p = a0 & 1, h0 = a0 >> 1;
        float* const pe = p ? accO : accE;          // plane of the even taps
        float* const po = p ? accE : accO;          // plane of the odd taps
        float2 ae[NZ / 2], ao[NZ / 2];
#pragma unroll
        for (int m = 0; m < NZ / 2; ++m) {
          const int q0 = lane + 64 * m, q1 = q0 + 32;
          ae[m].x = q0 < g.half ? pe[(h0 + q0) & kMask] : 0.0f;
          ao[m].x = q0 < g.half ? po[(h0 + q0 + p) & kMask] : 0.0f;
          ae[m].y = q1 < g.half ? pe[(h0 + q1) & kMask] : 0.0f;
          ao[m].y = q1 < g.half ? po[(h0 + q1 + p) & kMask] : 0.0f;
        }
#pragma unroll
        for (int m = 0; m < NZ / 2; ++m) {
          ae[m] = __ffma2_rn(R[m], wE2[m * 32 + lane], ae[m]);
          ao[m] = __ffma2_rn(neg2(I[m]), wO2[m * 32 + lane], ao[m]);
        }
#pragma unroll
        for (int m = 0; m < NZ / 2; ++m) {
          const int q0 = lane + 64 * m, q1 = q0 + 32;
          if (q0 < g.half) { pe[(h0 + q0) & kMask] = ae[m].x; po[(h0 + q0 + p) & kMask] = ao[m].x; }
          if (q1 < g.half) { pe[(h0 + q1) & kMask] = ae[m].y; po[(h0 + q1 + p) & kMask] = ao[m].y; }
        }
      }
      // bar.arrive orders this warp's prior shared-memory writes before the barrier completes (PTX producer/consumer idiom)
      if (s + 1 < n_seq) named_bar_arrive(1 + warp, 64);                      // release the owner of frame s+1

      // ---- samples [t*hop - off0, (t+1)*hop - off0) are final now (the utterance's last frame completes the rest)
      {
        const int n_out = (t == T - 1) ? g.off0 : g.hop;
        for (int r = lane; r < n_out; r += 32) {
          const int av = a0 + r;
          float* const cell = ((av & 1) ? accO : accE) + ((av >> 1) & kMask);
          const float val = *cell;
          *cell = 0.0f;
          const int i = i_start + r;
          if (own && i >= 0) {
            float inv;
            if (r < g.hop && t >= ND - 1) {
              inv = pw[r];
            } else {                                   // some overlapping frame does not exist: explicit window sum
              float ws = 0.0f;
              for (int d = 0, m = r; m < g.win && d <= t; ++d, m += g.hop) {
                const float wv = (m & 1) ? __ldg(tb.wO + (m >> 1)) : __ldg(tb.wE + (m >> 1));
                ws = fmaf(wv, wv, ws);
              }
              inv = ws > 1.17549435e-38f ? kInvN / ws : kInvN;   // librosa: divide only where wss > tiny
            }
            dst[i] = val * inv;
          }
        }
      }
    }  // frames of this warp

    if constexpr (SC) {
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
