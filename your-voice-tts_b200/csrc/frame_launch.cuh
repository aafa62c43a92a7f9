// Launch plumbing for the frame kernels (instantiated in frame_gl.cu / frame_synth.cu / frame_analysis.cu).
#pragma once
#include <atomic>
#include "frame_kernels.cuh"

namespace ttsa {

extern std::atomic<unsigned long long> g_launches;

// Opt every instantiation in to `smem_bytes` of dynamic shared memory and report the resident CTAs per SM of the
// Griffin-Lim iteration kernel.  Returns nullptr or an error string.
const char* configure_frame_kernels(size_t smem_bytes, int* ctas_per_sm);
const char* launch_frame_kernel(int mode, int src, int nz, bool sc, int grid, size_t smem_bytes, cudaStream_t st,
                                const Geo& g, const Tables& tb, const BatchDev& bd, const FrameArgs& a);

// per translation unit
const char* configure_gl(size_t smem_bytes, int* ctas_per_sm);
const char* configure_synth(size_t smem_bytes);
const char* configure_analysis(size_t smem_bytes);
const char* launch_gl(int src, int nz, bool sc, int grid, size_t smem, cudaStream_t st, const Geo&, const Tables&, const BatchDev&, const FrameArgs&);
// barrier-free warp-chain Griffin-Lim iteration (standard geometry class); smem layout in gl_chain.cuh
struct ChainSmem;
const char* configure_gl_chain(size_t smem_bytes, int* ctas_per_sm);
const char* launch_gl_chain(int src, bool sc, int grid, size_t smem, cudaStream_t st, const Geo&, const Tables&, const BatchDev&,
                            const FrameArgs&, const ChainSmem&, long long total_frames);
const char* launch_synth(int src, int nz, int grid, size_t smem, cudaStream_t st, const Geo&, const Tables&, const BatchDev&, const FrameArgs&);
const char* launch_analysis(int out, int nz, int grid, size_t smem, cudaStream_t st, const Geo&, const Tables&, const BatchDev&, const FrameArgs&);

template <class K>
inline const char* set_smem(K kernel, size_t smem_bytes) {
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
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

}  // namespace ttsa
