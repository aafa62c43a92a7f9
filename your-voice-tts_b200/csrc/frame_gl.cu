// Griffin-Lim iteration kernel instantiations (MODE_GL_ITER).
#include "frame_launch.cuh"
#include "gl_chain.cuh"

namespace ttsa {

const char* configure_gl(size_t smem_bytes, int* ctas_per_sm) {
  const char* e;
  if ((e = set_smem(frame_kernel<MODE_GL_ITER, SRC_MAG, 20, false>, smem_bytes))) return e;
  if ((e = set_smem(frame_kernel<MODE_GL_ITER, SRC_MAG, 20, true>, smem_bytes))) return e;
  if ((e = set_smem(frame_kernel<MODE_GL_ITER, SRC_MAG, 32, false>, smem_bytes))) return e;
  if ((e = set_smem(frame_kernel<MODE_GL_ITER, SRC_MAG, 32, true>, smem_bytes))) return e;
  if ((e = set_smem(frame_kernel<MODE_GL_ITER, SRC_NORM_DB, 20, false>, smem_bytes))) return e;
  if ((e = set_smem(frame_kernel<MODE_GL_ITER, SRC_NORM_DB, 20, true>, smem_bytes))) return e;
  if ((e = set_smem(frame_kernel<MODE_GL_ITER, SRC_NORM_DB, 32, false>, smem_bytes))) return e;
  if ((e = set_smem(frame_kernel<MODE_GL_ITER, SRC_NORM_DB, 32, true>, smem_bytes))) return e;
  int occ = 0;
  cudaError_t ce = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, frame_kernel<MODE_GL_ITER, SRC_MAG, 20, false>,
                                                                 kThreads, smem_bytes);
  if (ce != cudaSuccess) return cudaGetErrorString(ce);
  *ctas_per_sm = occ;
  return nullptr;
}

const char* launch_gl(int src, int nz, bool sc, int grid, size_t smem, cudaStream_t st, const Geo& g, const Tables& tb,
                      const BatchDev& bd, const FrameArgs& a) {
  if (src == SRC_MAG) {
    if (nz == 20) { if (sc) TTSA_LAUNCH((frame_kernel<MODE_GL_ITER, SRC_MAG, 20, true>)); else TTSA_LAUNCH((frame_kernel<MODE_GL_ITER, SRC_MAG, 20, false>)); }
    else          { if (sc) TTSA_LAUNCH((frame_kernel<MODE_GL_ITER, SRC_MAG, 32, true>)); else TTSA_LAUNCH((frame_kernel<MODE_GL_ITER, SRC_MAG, 32, false>)); }
  } else {
    if (nz == 20) { if (sc) TTSA_LAUNCH((frame_kernel<MODE_GL_ITER, SRC_NORM_DB, 20, true>)); else TTSA_LAUNCH((frame_kernel<MODE_GL_ITER, SRC_NORM_DB, 20, false>)); }
    else          { if (sc) TTSA_LAUNCH((frame_kernel<MODE_GL_ITER, SRC_NORM_DB, 32, true>)); else TTSA_LAUNCH((frame_kernel<MODE_GL_ITER, SRC_NORM_DB, 32, false>)); }
  }
}

const char* configure_gl_chain(size_t smem_bytes, int* ctas_per_sm) {
  const char* e;
  if ((e = set_smem(gl_chain_kernel<SRC_MAG, false>, smem_bytes))) return e;
  if ((e = set_smem(gl_chain_kernel<SRC_MAG, true>, smem_bytes))) return e;
  if ((e = set_smem(gl_chain_kernel<SRC_NORM_DB, false>, smem_bytes))) return e;
  if ((e = set_smem(gl_chain_kernel<SRC_NORM_DB, true>, smem_bytes))) return e;
  int occ = 0;
  cudaError_t ce = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, gl_chain_kernel<SRC_MAG, false>, kThreads, smem_bytes);
  if (ce != cudaSuccess) return cudaGetErrorString(ce);
  *ctas_per_sm = occ;
  return nullptr;
}

#define TTSA_LAUNCH_CHAIN(KERNEL)                                                  \
  do {                                                                             \
    KERNEL<<<grid, kThreads, smem, st>>>(g, tb, bd, a, sm, total_frames);          \
    g_launches += 1;                                                               \
    cudaError_t e_ = cudaGetLastError();                                           \
    return e_ == cudaSuccess ? nullptr : cudaGetErrorString(e_);                   \
  } while (0)

const char* launch_gl_chain(int src, bool sc, int grid, size_t smem, cudaStream_t st, const Geo& g, const Tables& tb,
                            const BatchDev& bd, const FrameArgs& a, const ChainSmem& sm, long long total_frames) {
  if (src == SRC_MAG) { if (sc) TTSA_LAUNCH_CHAIN((gl_chain_kernel<SRC_MAG, true>)); else TTSA_LAUNCH_CHAIN((gl_chain_kernel<SRC_MAG, false>)); }
  else                { if (sc) TTSA_LAUNCH_CHAIN((gl_chain_kernel<SRC_NORM_DB, true>)); else TTSA_LAUNCH_CHAIN((gl_chain_kernel<SRC_NORM_DB, false>)); }
}

}  // namespace ttsa
