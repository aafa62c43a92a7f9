// Compile-time helpers of the in-register 32-point FFT (fft32p.cuh): unrolled loop, bit reversal, twiddle constants.
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

}  // namespace ttsa
