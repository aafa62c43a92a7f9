// Griffin-Lim iteration kernel instantiations (MODE_GL_ITER).
#include "frame_launch.cuh"

namespace ttsa {

const char* configure_gl_mom(size_t smem_bytes);
const char* launch_gl_mom(int src, int nz, bool sc, bool fixed, int grid, size_t smem, cudaStream_t st, const Geo& g,
                          const Tables& tb, const BatchDev& bd, const FrameArgs& a);

const char* configure_gl(size_t smem_bytes, int* ctas_per_sm) {
  const char* e;
  if ((e = configure_variants<MODE_GL_ITER, SRC_MAG, false>(smem_bytes))) return e;
  if ((e = configure_variants<MODE_GL_ITER, SRC_MAG, true>(smem_bytes))) return e;
  if ((e = configure_variants<MODE_GL_ITER, SRC_NORM_DB, false>(smem_bytes))) return e;
  if ((e = configure_variants<MODE_GL_ITER, SRC_NORM_DB, true>(smem_bytes))) return e;
  if ((e = configure_gl_mom(smem_bytes))) return e;
  if ((e = configure_fine<SRC_MAG>(smem_bytes))) return e;
  if ((e = configure_fine<SRC_NORM_DB>(smem_bytes))) return e;
  int occ = 0;
  cudaError_t ce = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, frame_kernel<MODE_GL_ITER, SRC_MAG, 20, false>,
                                                                 kThreads, smem_bytes);
  if (ce != cudaSuccess) return cudaGetErrorString(ce);
  *ctas_per_sm = occ;
  return nullptr;
}

const char* launch_gl(int src, int nz, bool sc, bool fixed, bool mom, int grid, size_t smem, cudaStream_t st, const Geo& g,
                      const Tables& tb, const BatchDev& bd, const FrameArgs& a, bool fine) {
  if (mom) return launch_gl_mom(src, nz, sc, fixed, grid, smem, st, g, tb, bd, a);
  if (fine && !sc && nz == 20)
    return src == SRC_MAG ? launch_fine<SRC_MAG>(fixed, grid, smem, st, g, tb, bd, a) : launch_fine<SRC_NORM_DB>(fixed, grid, smem, st, g, tb, bd, a);
  if (src == SRC_MAG) return sc ? launch_variant<MODE_GL_ITER, SRC_MAG, true>(nz, fixed, grid, smem, st, g, tb, bd, a)
                                : launch_variant<MODE_GL_ITER, SRC_MAG, false>(nz, fixed, grid, smem, st, g, tb, bd, a);
  return sc ? launch_variant<MODE_GL_ITER, SRC_NORM_DB, true>(nz, fixed, grid, smem, st, g, tb, bd, a)
            : launch_variant<MODE_GL_ITER, SRC_NORM_DB, false>(nz, fixed, grid, smem, st, g, tb, bd, a);
}

}  // namespace ttsa
