// Griffin-Lim iteration kernel with momentum input (fast Griffin-Lim, opt-in): the transform sees
// wav_in - beta * wav_prev.  Separate translation unit so that the default kernels build in parallel with these.
#include "frame_launch.cuh"

namespace ttsa {

const char* configure_gl_mom(size_t smem_bytes) {
  const char* e;
  if ((e = configure_variants<MODE_GL_ITER, SRC_MAG, false, true>(smem_bytes))) return e;
  if ((e = configure_variants<MODE_GL_ITER, SRC_MAG, true, true>(smem_bytes))) return e;
  if ((e = configure_variants<MODE_GL_ITER, SRC_NORM_DB, false, true>(smem_bytes))) return e;
  return configure_variants<MODE_GL_ITER, SRC_NORM_DB, true, true>(smem_bytes);
}

const char* launch_gl_mom(int src, int nz, bool sc, bool fixed, int grid, size_t smem, cudaStream_t st, const Geo& g,
                          const Tables& tb, const BatchDev& bd, const FrameArgs& a) {
  if (src == SRC_MAG) return sc ? launch_variant<MODE_GL_ITER, SRC_MAG, true, true>(nz, fixed, grid, smem, st, g, tb, bd, a)
                                : launch_variant<MODE_GL_ITER, SRC_MAG, false, true>(nz, fixed, grid, smem, st, g, tb, bd, a);
  return sc ? launch_variant<MODE_GL_ITER, SRC_NORM_DB, true, true>(nz, fixed, grid, smem, st, g, tb, bd, a)
            : launch_variant<MODE_GL_ITER, SRC_NORM_DB, false, true>(nz, fixed, grid, smem, st, g, tb, bd, a);
}

}  // namespace ttsa
