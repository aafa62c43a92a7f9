// Per-warp building blocks shared by the frame kernels: the 1024-point transform (32 x 32, packed fp32 pairs, one
// shared-memory exchange) and the Griffin-Lim per-bin step (phase projection on conjugate pairs).
#pragma once
#include "frame_kernels.cuh"

namespace ttsa {

// One 32-point pass of the 1024-point forward transform of a frame held by a warp (element n2 of lane L =
// z[L + 32 n2]  ->  after both passes element k1 of lane L = Z[32 k1 + L]).  Pass 0 ends with the twiddle multiply and
// the shared-memory exchange through the warp's buffer `buf` (kBufFloats); after it the buffer is idle until the next
// transform, which is where callers start asynchronous copies into it.  Callers roll the two passes (and the two
// transforms of a Griffin-Lim iteration) into loops so that this body exists once in the instruction stream.
__device__ __forceinline__ void transform_pass(float2 (&R)[16], float2 (&I)[16], float* buf, const float4* tw4, int lane,
                                               bool first_pass) {
  fft32p(R, I);
  if (first_pass) {
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
    __syncwarp();                                       // every lane has its row: the buffer is free
  }
}

// Griffin-Lim per-bin step (utils/audio.py:187-188): X -> Y = |S| X/|X| on the real-FFT bins, expressed on the packed
// 1024-point spectrum.  In: element k1 of lane L = Z[32 k1 + L].  Out: conj(Z') ready for the forward transform that
// implements the inverse.  Bin k pairs with 1024 - k, held by lane (32 - L) & 31 in register 31 - k1 (lane 0: its own
// register 32 - k1); each lane processes the 16 pairs whose first member is its own register k1 < 16, two pairs
// (k1 = 2m, 2m+1) per packed instruction.  srow = this frame's |S| (or normalised dB) row in shared memory.
template <int SRC, bool SC>
__device__ __forceinline__ void gl_bin_step(float2 (&R)[16], float2 (&I)[16], const float* srow, const float4* g4,
                                            const Geo& g, int lane, bool count_sc, float& sc_num, float& sc_den) {
  constexpr float kTiny = 1e-37f;
  const int partner = (32 - lane) & 31;
  const bool l0 = lane == 0;
  float2 BR[8], BI[8];
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
  float2 SR[8], SI[8];
  static_for<0, 8>([&](auto mc) {
    constexpr int m = decltype(mc)::value;
    const int k0 = 64 * m + lane;                      // bins k0 (k1 = 2m) and k0 + 32 (k1 = 2m+1); partners 1024 - k
    const float4 gq = g4[m * 32 + lane];
    const float2 GX = make_float2(gq.x, gq.y), GY = make_float2(gq.z, gq.w);
    const float2 Sk = make_float2(spec_to_mag<SRC>(srow[k0], g), spec_to_mag<SRC>(srow[k0 + 32], g));
    const float2 Sp = make_float2(spec_to_mag<SRC>(srow[1024 - k0], g), spec_to_mag<SRC>(srow[992 - k0], g));
    const float2 E2R = __fadd2_rn(R[m], BR[m]), E2I = __fadd2_rn(I[m], neg2(BI[m]));
    const float2 D2R = __fadd2_rn(R[m], neg2(BR[m])), D2I = __fadd2_rn(I[m], BI[m]);
    // 2 X[k] = E2 + G D2 ;  2 X[1024-k] = conj(E2 - G D2) = conj(2 E2 - 2 X[k])
    float2 XkR = __ffma2_rn(GX, D2R, E2R);
    XkR = __ffma2_rn(neg2(GY), D2I, XkR);
    float2 XkI = __ffma2_rn(GX, D2I, E2I);
    XkI = __ffma2_rn(GY, D2R, XkI);
    const float2 XpR = __ffma2_rn(E2R, splat(2.0f), neg2(XkR));
    const float2 XpI = __ffma2_rn(E2I, splat(-2.0f), XkI);
    const float2 mk = __ffma2_rn(XkI, XkI, __fmul2_rn(XkR, XkR));
    const float2 mp = __ffma2_rn(XpI, XpI, __fmul2_rn(XpR, XpR));
    const float2 ik = make_float2(rsqrt_fast(fmaxf(mk.x, kTiny)), rsqrt_fast(fmaxf(mk.y, kTiny)));
    const float2 ip = make_float2(rsqrt_fast(fmaxf(mp.x, kTiny)), rsqrt_fast(fmaxf(mp.y, kTiny)));
    const float2 fk = __fmul2_rn(Sk, ik), fp = __fmul2_rn(Sp, ip);
    // Y = S X/|X|;  np.angle(0) = 0  ->  Y = S  (the imaginary part is 0 * finite = 0 already)
    float2 YkR = __fmul2_rn(XkR, fk);
    const float2 YkI = __fmul2_rn(XkI, fk);
    float2 YpR = __fmul2_rn(XpR, fp);
    const float2 YpI = __fmul2_rn(XpI, fp);
    YkR.x = mk.x > kTiny ? YkR.x : Sk.x; YkR.y = mk.y > kTiny ? YkR.y : Sk.y;
    YpR.x = mp.x > kTiny ? YpR.x : Sp.x; YpR.y = mp.y > kTiny ? YpR.y : Sp.y;
    if (SC && count_sc) {
      const float2 dk = __ffma2_rn(__fmul2_rn(mk, ik), splat(0.5f), neg2(Sk));   // |X| - S
      const float2 dp = __ffma2_rn(__fmul2_rn(mp, ip), splat(0.5f), neg2(Sp));
      sc_num += dk.x * dk.x + dk.y * dk.y + dp.x * dp.x + dp.y * dp.y;
      sc_den += Sk.x * Sk.x + Sk.y * Sk.y + Sp.x * Sp.x + Sp.y * Sp.y;
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
  float2 z512 = make_float2(0.0f, 0.0f);
  if (l0) {   // k = 512 (self-paired, register 16 = R[8].x): X = conj(Z[512]), Z'2 = 2 conj(Y), conj(Z'2) = 2 Y
    const float S5 = spec_to_mag<SRC>(srow[512], g);
    const float2 X = make_float2(R[8].x, -I[8].x);
    const float m = X.x * X.x + X.y * X.y;
    const float im = rsqrt_fast(fmaxf(m, kTiny));
    const float f = S5 * im;
    z512 = make_float2(2.0f * (m > kTiny ? X.x * f : S5), 2.0f * X.y * f);
    if (SC && count_sc) {
      const float d = m * im - S5;         // |X| = |Z[512]| (no factor 2 here)
      sc_num += d * d;
      sc_den += S5 * S5;
    }
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

}  // namespace ttsa
