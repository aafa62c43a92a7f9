// In-register 32-point complex FFT (radix-2 decimation in time, fully unrolled, compile-time twiddles).
//
// One thread transforms 32 complex values held in 64 registers.  Natural-order input and output: the DIT
// bit reversal is a compile-time register renaming.  Non-trivial butterflies use the 6-instruction FMA form
// (s = a + w*b with 4 FMAs, d = 2a - s with 2 FMAs).  NZ < 32 declares inputs [NZ, 32) to be identically zero
// (the window covers only ceil(win/2) of the 1024 packed samples), which removes the first-stage butterflies
// whose second operand is zero; unused outputs are removed by dead-code elimination at the call site.
#pragma once
#include <cuda_runtime.h>
#include <type_traits>

namespace ttsa {

template <int I, int N, class F>
__device__ __forceinline__ void static_for(F&& f) {
  if constexpr (I < N) {
    f(std::integral_constant<int, I>{});
    static_for<I + 1, N>(f);
  }
}

__host__ __device__ constexpr int brev5(int i) {
  return ((i & 1) << 4) | ((i & 2) << 2) | (i & 4) | ((i & 8) >> 2) | ((i & 16) >> 4);
}

// cos / sin of 2*pi*k/32, k = 0..15
struct Tw32 {
  static constexpr float c[16] = {1.0f, 0.98078528040323044913f, 0.92387953251128675613f, 0.83146961230254523708f,
                                  0.70710678118654752440f, 0.55557023301960222474f, 0.38268343236508977173f,
                                  0.19509032201612826785f, 0.0f, -0.19509032201612826785f, -0.38268343236508977173f,
                                  -0.55557023301960222474f, -0.70710678118654752440f, -0.83146961230254523708f,
                                  -0.92387953251128675613f, -0.98078528040323044913f};
  static constexpr float s[16] = {0.0f, 0.19509032201612826785f, 0.38268343236508977173f, 0.55557023301960222474f,
                                  0.70710678118654752440f, 0.83146961230254523708f, 0.92387953251128675613f,
                                  0.98078528040323044913f, 1.0f, 0.98078528040323044913f, 0.92387953251128675613f,
                                  0.83146961230254523708f, 0.70710678118654752440f, 0.55557023301960222474f,
                                  0.38268343236508977173f, 0.19509032201612826785f};
};

// (a, b) <- (a + w b, a - w b),  w = exp(-/+ j 2 pi K / 32)
template <int K, bool INV>
__device__ __forceinline__ void bfly(float2& a, float2& b) {
  if constexpr (K == 0) {
    const float2 s = make_float2(a.x + b.x, a.y + b.y);
    const float2 d = make_float2(a.x - b.x, a.y - b.y);
    a = s; b = d;
  } else if constexpr (K == 8) {
    // forward w = -j: w b = (b.y, -b.x); inverse w = +j: w b = (-b.y, b.x)
    const float tx = INV ? -b.y : b.y;
    const float ty = INV ? b.x : -b.x;
    const float2 s = make_float2(a.x + tx, a.y + ty);
    const float2 d = make_float2(a.x - tx, a.y - ty);
    a = s; b = d;
  } else {
    constexpr float wr = Tw32::c[K];
    constexpr float wi = INV ? Tw32::s[K] : -Tw32::s[K];
    float sx = fmaf(wr, b.x, a.x);
    sx = fmaf(-wi, b.y, sx);
    float sy = fmaf(wr, b.y, a.y);
    sy = fmaf(wi, b.x, sy);
    b.x = fmaf(2.0f, a.x, -sx);
    b.y = fmaf(2.0f, a.y, -sy);
    a.x = sx; a.y = sy;
  }
}

template <bool INV, int NZ = 32>
__device__ __forceinline__ void fft32(float2 (&v)[32]) {
  float2 a[32];
  static_for<0, 32>([&](auto ic) {
    constexpr int i = decltype(ic)::value;
    a[i] = v[brev5(i)];
  });
  // stage h = 1: pairs (a[2m], a[2m+1]) = (v[r], v[r+16]), r = brev5(2m) < 16
  static_for<0, 16>([&](auto mc) {
    constexpr int m = decltype(mc)::value;
    constexpr int r = brev5(2 * m);
    if constexpr (r + 16 >= NZ) {
      a[2 * m + 1] = a[2 * m];       // second operand identically zero: sum = difference = a
    } else {
      bfly<0, INV>(a[2 * m], a[2 * m + 1]);
    }
  });
  // stages h = 2, 4, 8, 16
  static_for<1, 5>([&](auto stc) {
    constexpr int h = 1 << decltype(stc)::value;
    static_for<0, 16>([&](auto pc) {
      constexpr int p = decltype(pc)::value;
      constexpr int grp = p / h, j = p % h;
      constexpr int i0 = grp * 2 * h + j;
      bfly<j * (16 / h), INV>(a[i0], a[i0 + h]);
    });
  });
  static_for<0, 32>([&](auto ic) {
    constexpr int i = decltype(ic)::value;
    v[i] = a[i];
  });
}

}  // namespace ttsa
