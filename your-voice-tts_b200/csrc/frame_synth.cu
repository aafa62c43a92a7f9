// Spectrum -> waveform kernel instantiations (MODE_SYNTH: initial Griffin-Lim estimate and _istft).
#include "frame_launch.cuh"

namespace ttsa {

const char* configure_synth(size_t smem_bytes) {
  const char* e;
  if ((e = configure_variants<MODE_SYNTH, SRC_MAG, false>(smem_bytes))) return e;
  if ((e = configure_variants<MODE_SYNTH, SRC_NORM_DB, false>(smem_bytes))) return e;
  return configure_variants<MODE_SYNTH, SRC_COMPLEX, false>(smem_bytes);
}

const char* launch_synth(int src, int nz, bool fixed, int grid, size_t smem, cudaStream_t st, const Geo& g, const Tables& tb,
                         const BatchDev& bd, const FrameArgs& a) {
  if (src == SRC_MAG) return launch_variant<MODE_SYNTH, SRC_MAG, false>(nz, fixed, grid, smem, st, g, tb, bd, a);
  if (src == SRC_NORM_DB) return launch_variant<MODE_SYNTH, SRC_NORM_DB, false>(nz, fixed, grid, smem, st, g, tb, bd, a);
  return launch_variant<MODE_SYNTH, SRC_COMPLEX, false>(nz, fixed, grid, smem, st, g, tb, bd, a);
}

}  // namespace ttsa
