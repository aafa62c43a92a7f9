// Launch plumbing for the frame kernels (instantiated in frame_gl.cu / frame_synth.cu / frame_analysis.cu).
#pragma once
#include <atomic>
#include "frame_kernels.cuh"

namespace ttsa {

extern std::atomic<unsigned long long> g_launches;

// Opt every instantiation in to `smem_bytes` of dynamic shared memory and report the resident CTAs per SM of the
// Griffin-Lim iteration kernel.  Returns nullptr or an error string.
const char* configure_frame_kernels(size_t smem_bytes, int* ctas_per_sm);
const char* launch_frame_kernel(int mode, int src, int nz, bool sc, bool fixed, bool mom, int grid, size_t smem_bytes, cudaStream_t st,
                                const Geo& g, const Tables& tb, const BatchDev& bd, const FrameArgs& a, bool fine = false);

// per translation unit
const char* configure_gl(size_t smem_bytes, int* ctas_per_sm);
const char* configure_synth(size_t smem_bytes);
const char* configure_analysis(size_t smem_bytes);
const char* launch_gl(int src, int nz, bool sc, bool fixed, bool mom, int grid, size_t smem, cudaStream_t st, const Geo&, const Tables&, const BatchDev&, const FrameArgs&, bool fine = false);
const char* launch_synth(int src, int nz, bool fixed, int grid, size_t smem, cudaStream_t st, const Geo&, const Tables&, const BatchDev&, const FrameArgs&);
const char* launch_analysis(int out, int nz, bool fixed, int grid, size_t smem, cudaStream_t st, const Geo&, const Tables&, const BatchDev&, const FrameArgs&);

// warp-stream Griffin-Lim iteration kernel (gl_stream.cu)
bool gl_stream_supported(int hop, int win);
const char* configure_gl_stream();
const char* launch_gl_stream(int src, bool sc, int hop, int win, int grid, cudaStream_t st, const Geo&, const Tables&, const BatchDev&,
                             const WpsDev&, const FrameArgs&);

// warp-stream feature extraction kernel (feat_stream.cu)
bool feat_stream_supported(int hop, int win, int mel_smem_floats);
const char* configure_feat_stream();
const char* launch_feat_stream(int hop, int win, int grid, cudaStream_t st, const Geo&, const Tables&, const BatchDev&, const FrameArgs&,
                               int total_frames);

template <class K>
inline const char* set_smem(K kernel, size_t /*smem_bytes*/) {
  // The attribute belongs to the function, not to a plan: opt in to the device maximum (227 KB on sm_100) once, so that
  // plans with different geometries can share the kernels; each launch still requests only what its plan needs.
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}

// Frame kernels are launched with programmatic dependent launch: the next kernel on the stream may start its
// prologue (table loads into shared memory) while this one drains; every kernel executes griddepcontrol.wait before it
// touches global data produced by its predecessor.
#define TTSA_LAUNCH(KERNEL)                                                        \
  do {                                                                             \
    cudaLaunchConfig_t cfg_ = {};                                                  \
    cfg_.gridDim = dim3((unsigned)grid);                                           \
    cfg_.blockDim = dim3(kThreads);                                                \
    cfg_.dynamicSmemBytes = smem;                                                  \
    cfg_.stream = st;                                                              \
    cudaLaunchAttribute attr_[1];                                                  \
    attr_[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;              \
    attr_[0].val.programmaticStreamSerializationAllowed = 1;                       \
    cfg_.attrs = attr_;                                                            \
    cfg_.numAttrs = 1;                                                             \
    cudaError_t e_ = cudaLaunchKernelEx(&cfg_, KERNEL, g, tb, bd, a);              \
    g_launches += 1;                                                               \
    if (e_ == cudaSuccess) e_ = cudaGetLastError();                                \
    return e_ == cudaSuccess ? nullptr : cudaGetErrorString(e_);                   \
  } while (0)


// The shipped (hop, win) geometries (22.05 kHz config.json, 16 kHz config_tacotron_de.json, 24 kHz config_libritts.json)
// get class-20 kernels with the geometry as template constants; everything else runs the run-time-geometry kernels.
#define TTSA_FIXED_GEOS(X) X(275, 1102) X(200, 800) X(300, 1200)

// The opt-in momentum variant (fast Griffin-Lim, not in the reference) is instantiated for run-time geometry only.
template <int MODE, int SRC, bool SC, bool MOM = false>
inline const char* configure_variants(size_t smem_bytes) {
  const char* e;
  if ((e = set_smem(frame_kernel<MODE, SRC, 20, SC, 0, 0, MOM>, smem_bytes))) return e;
  if ((e = set_smem(frame_kernel<MODE, SRC, 32, SC, 0, 0, MOM>, smem_bytes))) return e;
  if constexpr (!MOM) {
#define TTSA_X(H, W) if ((e = set_smem(frame_kernel<MODE, SRC, 20, SC, H, W, MOM>, smem_bytes))) return e;
    TTSA_FIXED_GEOS(TTSA_X)
#undef TTSA_X
  }
  return nullptr;
}

// fine-segment form of the Griffin-Lim iteration (small batches; no spectral-convergence sums, no momentum)
template <int SRC>
inline const char* configure_fine(size_t smem_bytes) {
  const char* e;
  if ((e = set_smem(frame_kernel<MODE_GL_ITER, SRC, 20, false, 0, 0, false, true>, smem_bytes))) return e;
#define TTSA_X(H, W) if ((e = set_smem(frame_kernel<MODE_GL_ITER, SRC, 20, false, H, W, false, true>, smem_bytes))) return e;
  TTSA_FIXED_GEOS(TTSA_X)
#undef TTSA_X
  return nullptr;
}
template <int SRC>
inline const char* launch_fine(bool fixed, int grid, size_t smem, cudaStream_t st, const Geo& g, const Tables& tb,
                               const BatchDev& bd, const FrameArgs& a) {
  if (fixed) {
#define TTSA_X(H, W) if (g.ly.hop == H && g.ly.win == W) TTSA_LAUNCH((frame_kernel<MODE_GL_ITER, SRC, 20, false, H, W, false, true>));
    TTSA_FIXED_GEOS(TTSA_X)
#undef TTSA_X
  }
  TTSA_LAUNCH((frame_kernel<MODE_GL_ITER, SRC, 20, false, 0, 0, false, true>));
}

template <int MODE, int SRC, bool SC, bool MOM = false>
inline const char* launch_variant(int nz, bool fixed, int grid, size_t smem, cudaStream_t st, const Geo& g, const Tables& tb,
                                  const BatchDev& bd, const FrameArgs& a) {
  if (nz != 20) TTSA_LAUNCH((frame_kernel<MODE, SRC, 32, SC, 0, 0, MOM>));
  if constexpr (!MOM) {
    if (fixed) {
#define TTSA_X(H, W) if (g.ly.hop == H && g.ly.win == W) TTSA_LAUNCH((frame_kernel<MODE, SRC, 20, SC, H, W, MOM>));
      TTSA_FIXED_GEOS(TTSA_X)
#undef TTSA_X
    }
  }
  TTSA_LAUNCH((frame_kernel<MODE, SRC, 20, SC, 0, 0, MOM>));
}

}  // namespace ttsa
