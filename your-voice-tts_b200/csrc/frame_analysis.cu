// Waveform -> spectrum kernel instantiations (MODE_ANALYSIS: _stft, spectrogram, melspectrogram) and the dispatcher.
#include "frame_launch.cuh"

namespace ttsa {

const char* configure_analysis(size_t smem_bytes) {
  const char* e;
  if ((e = configure_variants<MODE_ANALYSIS, OUT_COMPLEX, false>(smem_bytes))) return e;
  return configure_variants<MODE_ANALYSIS, OUT_FEATURES, false>(smem_bytes);
}

const char* launch_analysis(int out, int nz, bool fixed, int grid, size_t smem, cudaStream_t st, const Geo& g, const Tables& tb,
                            const BatchDev& bd, const FrameArgs& a) {
  if (out == OUT_COMPLEX) return launch_variant<MODE_ANALYSIS, OUT_COMPLEX, false>(nz, fixed, grid, smem, st, g, tb, bd, a);
  return launch_variant<MODE_ANALYSIS, OUT_FEATURES, false>(nz, fixed, grid, smem, st, g, tb, bd, a);
}

const char* configure_frame_kernels(size_t smem_bytes, int* ctas_per_sm) {
  const char* e;
  if ((e = configure_gl(smem_bytes, ctas_per_sm))) return e;
  if ((e = configure_synth(smem_bytes))) return e;
  return configure_analysis(smem_bytes);
}

const char* launch_frame_kernel(int mode, int src, int nz, bool sc, bool fixed, bool mom, int grid, size_t smem_bytes, cudaStream_t st,
                                const Geo& g, const Tables& tb, const BatchDev& bd, const FrameArgs& a, bool fine) {
  if (mode == MODE_GL_ITER) return launch_gl(src, nz, sc, fixed, mom, grid, smem_bytes, st, g, tb, bd, a, fine);
  if (mode == MODE_SYNTH) return launch_synth(src, nz, fixed, grid, smem_bytes, st, g, tb, bd, a);
  return launch_analysis(src, nz, fixed, grid, smem_bytes, st, g, tb, bd, a);
}

}  // namespace ttsa
