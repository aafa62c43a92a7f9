// Feature extraction kernel, warp-stream form (feat_stream.cuh): instantiations for the shipped geometries.
#include "frame_launch.cuh"
#include "feat_stream.cuh"

namespace ttsa {

template <int HOP, int WIN>
static const char* launch_one(int grid, cudaStream_t st, const Geo& g, const Tables& tb, const BatchDev& bd, const FrameArgs& a,
                              int total_frames) {
  if constexpr (!FeatGeo<HOP, WIN>::kFits) return "feat_stream: geometry does not fit shared memory";
  else {
    const size_t mel_floats = (size_t)feat_mel_floats(g);
    const size_t smem = ((size_t)FeatGeo<HOP, WIN>::sm_mel + (a.mel_out != nullptr ? mel_floats : 0)) * 4;
    if (a.lin_out != nullptr && a.mel_out != nullptr) feat_stream_kernel<HOP, WIN, true, true><<<grid, kWpsThreads, smem, st>>>(g, tb, bd, a, total_frames);
    else if (a.lin_out != nullptr) feat_stream_kernel<HOP, WIN, true, false><<<grid, kWpsThreads, smem, st>>>(g, tb, bd, a, total_frames);
    else feat_stream_kernel<HOP, WIN, false, true><<<grid, kWpsThreads, smem, st>>>(g, tb, bd, a, total_frames);
    g_launches += 1;
    const cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
  }
}

bool feat_stream_supported(int hop, int win, int mel_smem_floats) {
#define TTSA_X(H, W) if (hop == H && win == W) return FeatGeo<H, W>::kFits && ((size_t)FeatGeo<H, W>::sm_mel + mel_smem_floats) * 4 <= 227 * 1024;
  TTSA_FIXED_GEOS(TTSA_X)
#undef TTSA_X
  return false;
}

const char* configure_feat_stream() {
  const char* e;
#define TTSA_X(H, W) if constexpr (FeatGeo<H, W>::kFits) { \
    if ((e = set_smem(feat_stream_kernel<H, W, true, true>, 0))) return e; \
    if ((e = set_smem(feat_stream_kernel<H, W, true, false>, 0))) return e; \
    if ((e = set_smem(feat_stream_kernel<H, W, false, true>, 0))) return e; }
  TTSA_FIXED_GEOS(TTSA_X)
#undef TTSA_X
  return nullptr;
}

const char* launch_feat_stream(int hop, int win, int grid, cudaStream_t st, const Geo& g, const Tables& tb, const BatchDev& bd,
                               const FrameArgs& a, int total_frames) {
#define TTSA_X(H, W) if (hop == H && win == W) return launch_one<H, W>(grid, st, g, tb, bd, a, total_frames);
  TTSA_FIXED_GEOS(TTSA_X)
#undef TTSA_X
  return "feat_stream: geometry not instantiated";
}

}  // namespace ttsa
