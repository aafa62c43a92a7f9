// In-register 32-point complex FFT on PACKED fp32 pairs (sm_100 FFMA2 / FADD2 / FMUL2).
//
// One thread transforms 32 complex values held as 16 + 16 float2 registers in structure-of-arrays form:
//     R[m] = (re[2m], re[2m+1]),  I[m] = (im[2m], im[2m+1]),   m = 0..15      (natural order in, natural order out)
// Radix-2 decimation in time.  The DIT bit reversal makes the two halves of R[m] the internal elements e and e + 16
// (e = 4-bit reversal of m), so stages 1-4 (spans 1, 2, 4, 8) run one packed instruction for the two butterflies
// (e, e+h) and (e+16, e+16+h), which share their twiddle; the last stage (span 16) pairs the two halves of one
// register and is done in scalar.  Non-trivial butterflies use the FMA form s = a + w b (4 FMA), d = 2a - s (2 FMA).
// Issue slots per transform: 148 packed + 92 scalar = 240 (388 for the scalar version in fft32.cuh); the FP32 lane
// work is the same -- the packed form frees issue slots for the load/store/shuffle instructions around it.
#pragma once
#include <cuda_runtime.h>
#include "fft32.cuh"

namespace ttsa {

__host__ __device__ constexpr int brev4(int i) { return ((i & 1) << 3) | ((i & 2) << 1) | ((i & 4) >> 1) | ((i & 8) >> 3); }

__device__ __forceinline__ float2 splat(float c) { return make_float2(c, c); }
__device__ __forceinline__ float2 neg2(float2 a) { return make_float2(-a.x, -a.y); }

// packed butterfly on (ar, ai), (br, bi):  a <- a + w b,  b <- a - w b,  w = exp(-j 2 pi K / 32)
template <int K>
__device__ __forceinline__ void bfly_p(float2& ar, float2& ai, float2& br, float2& bi) {
  if constexpr (K == 0) {
    const float2 sr = __fadd2_rn(ar, br), si = __fadd2_rn(ai, bi);
    br = __fadd2_rn(ar, neg2(br)); bi = __fadd2_rn(ai, neg2(bi));
    ar = sr; ai = si;
  } else if constexpr (K == 8) {      // w = -j:  w b = (b.im, -b.re)
    const float2 sr = __fadd2_rn(ar, bi), si = __fadd2_rn(ai, neg2(br));
    const float2 dr = __fadd2_rn(ar, neg2(bi)), di = __fadd2_rn(ai, br);
    ar = sr; ai = si; br = dr; bi = di;
  } else {
    constexpr float wr = Tw32::c[K];
    constexpr float wi = -Tw32::s[K];
    float2 sr = __ffma2_rn(br, splat(wr), ar);
    sr = __ffma2_rn(bi, splat(-wi), sr);
    float2 si = __ffma2_rn(bi, splat(wr), ai);
    si = __ffma2_rn(br, splat(wi), si);
    br = __ffma2_rn(ar, splat(2.0f), neg2(sr));
    bi = __ffma2_rn(ai, splat(2.0f), neg2(si));
    ar = sr; ai = si;
  }
}

// scalar butterfly for the last stage:  (a, b) <- (a + w b, a - w b),  w = exp(-j 2 pi K / 32)
template <int K>
__device__ __forceinline__ void bfly_s(float& ar, float& ai, float& br, float& bi) {
  if constexpr (K == 0) {
    const float sr = ar + br, si = ai + bi;
    br = ar - br; bi = ai - bi; ar = sr; ai = si;
  } else if constexpr (K == 8) {
    const float sr = ar + bi, si = ai - br, dr = ar - bi, di = ai + br;
    ar = sr; ai = si; br = dr; bi = di;
  } else {
    constexpr float wr = Tw32::c[K];
    constexpr float wi = -Tw32::s[K];
    float sr = fmaf(wr, br, ar);
    sr = fmaf(-wi, bi, sr);
    float si = fmaf(wr, bi, ai);
    si = fmaf(wi, br, si);
    br = fmaf(2.0f, ar, -sr);
    bi = fmaf(2.0f, ai, -si);
    ar = sr; ai = si;
  }
}

__device__ __forceinline__ void fft32p(float2 (&R)[16], float2 (&I)[16]) {
  // internal packed register e holds internal elements (e, e + 16) = natural (2m, 2m + 1), m = brev4(e)
  float2 pr[16], pi[16];
  static_for<0, 16>([&](auto ec) {
    constexpr int e = decltype(ec)::value;
    pr[e] = R[brev4(e)];
    pi[e] = I[brev4(e)];
  });
  static_for<0, 4>([&](auto stc) {
    constexpr int h = 1 << decltype(stc)::value;
    static_for<0, 8>([&](auto pc) {
      constexpr int p = decltype(pc)::value;
      constexpr int grp = p / h, j = p % h;
      constexpr int i0 = grp * 2 * h + j;
      bfly_p<j * (16 / h)>(pr[i0], pi[i0], pr[i0 + h], pi[i0 + h]);
    });
  });
  // last stage: internal e (.x) with e + 16 (.y), twiddle W_32^e; outputs X[e] and X[e + 16]
  float xr[32], xi[32];
  static_for<0, 16>([&](auto ec) {
    constexpr int e = decltype(ec)::value;
    float ar = pr[e].x, ai = pi[e].x, br = pr[e].y, bi = pi[e].y;
    bfly_s<e>(ar, ai, br, bi);
    xr[e] = ar; xi[e] = ai; xr[e + 16] = br; xi[e + 16] = bi;
  });
  static_for<0, 16>([&](auto mc) {
    constexpr int m = decltype(mc)::value;
    R[m] = make_float2(xr[2 * m], xr[2 * m + 1]);
    I[m] = make_float2(xi[2 * m], xi[2 * m + 1]);
  });
}

}  // namespace ttsa
