// Spectrum -> waveform kernel instantiations (MODE_SYNTH: initial Griffin-Lim estimate and _istft).
#include "frame_launch.cuh"

namespace ttsa {

const char* configure_synth(size_t smem_bytes) {
  const char* e;
  if ((e = set_smem(frame_kernel<MODE_SYNTH, SRC_MAG, 20, false>, smem_bytes))) return e;
  if ((e = set_smem(frame_kernel<MODE_SYNTH, SRC_MAG, 32, false>, smem_bytes))) return e;
  if ((e = set_smem(frame_kernel<MODE_SYNTH, SRC_NORM_DB, 20, false>, smem_bytes))) return e;
  if ((e = set_smem(frame_kernel<MODE_SYNTH, SRC_NORM_DB, 32, false>, smem_bytes))) return e;
  if ((e = set_smem(frame_kernel<MODE_SYNTH, SRC_COMPLEX, 20, false>, smem_bytes))) return e;
  if ((e = set_smem(frame_kernel<MODE_SYNTH, SRC_COMPLEX, 32, false>, smem_bytes))) return e;
  return nullptr;
}

const char* launch_synth(int src, int nz, int grid, size_t smem, cudaStream_t st, const Geo& g, const Tables& tb,
                         const BatchDev& bd, const FrameArgs& a) {
  if (src == SRC_MAG)          { if (nz == 20) TTSA_LAUNCH((frame_kernel<MODE_SYNTH, SRC_MAG, 20, false>)); else TTSA_LAUNCH((frame_kernel<MODE_SYNTH, SRC_MAG, 32, false>)); }
  else if (src == SRC_NORM_DB) { if (nz == 20) TTSA_LAUNCH((frame_kernel<MODE_SYNTH, SRC_NORM_DB, 20, false>)); else TTSA_LAUNCH((frame_kernel<MODE_SYNTH, SRC_NORM_DB, 32, false>)); }
  else                         { if (nz == 20) TTSA_LAUNCH((frame_kernel<MODE_SYNTH, SRC_COMPLEX, 20, false>)); else TTSA_LAUNCH((frame_kernel<MODE_SYNTH, SRC_COMPLEX, 32, false>)); }
}

}  // namespace ttsa
