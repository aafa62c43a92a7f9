// Griffin-Lim iteration kernel, warp-stream form (gl_stream.cuh): instantiations for the shipped geometries.
#include "frame_launch.cuh"
#include "gl_stream.cuh"

namespace ttsa {

template <int SRC, bool SC, int HOP, int WIN, bool FUSE>
static const char* launch_one(int grid, cudaStream_t st, const Geo& g, const Tables& tb, const BatchDev& bd, const WpsDev& wp,
                              const FrameArgs& a) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(kWpsThreads);
  cfg.dynamicSmemBytes = (size_t)WpsGeo<HOP, WIN>::sm_total * 4;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, gl_stream_kernel<SRC, SC, HOP, WIN, FUSE>, g, tb, bd, wp, a);
  g_launches += 1;
  if (e == cudaSuccess) e = cudaGetLastError();
  return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}

template <int HOP, int WIN>
static const char* configure_geo() {
  if constexpr (!WpsGeo<HOP, WIN>::kFits) return nullptr;
  const char* e;
  if ((e = set_smem(gl_stream_kernel<SRC_MAG, false, HOP, WIN, false>, 0))) return e;
  if ((e = set_smem(gl_stream_kernel<SRC_MAG, true, HOP, WIN, false>, 0))) return e;
  if ((e = set_smem(gl_stream_kernel<SRC_NORM_DB, false, HOP, WIN, false>, 0))) return e;
  if ((e = set_smem(gl_stream_kernel<SRC_NORM_DB, true, HOP, WIN, false>, 0))) return e;
  if ((e = set_smem(gl_stream_kernel<SRC_MAG, false, HOP, WIN, true>, 0))) return e;
  if ((e = set_smem(gl_stream_kernel<SRC_MAG, true, HOP, WIN, true>, 0))) return e;
  if ((e = set_smem(gl_stream_kernel<SRC_NORM_DB, false, HOP, WIN, true>, 0))) return e;
  return set_smem(gl_stream_kernel<SRC_NORM_DB, true, HOP, WIN, true>, 0);
}

// several iterations per launch (FrameArgs::wps_iters > 1) run the FUSE instantiation; the one-iteration kernel has the
// iteration loop compiled out
template <int HOP, int WIN, bool FUSE>
static const char* launch_geo_f(int src, bool sc, int grid, cudaStream_t st, const Geo& g, const Tables& tb, const BatchDev& bd,
                                const WpsDev& wp, const FrameArgs& a) {
  if constexpr (!WpsGeo<HOP, WIN>::kFits) return "gl_stream: geometry does not fit shared memory";
  else
  if (src == SRC_MAG) return sc ? launch_one<SRC_MAG, true, HOP, WIN, FUSE>(grid, st, g, tb, bd, wp, a)
                                : launch_one<SRC_MAG, false, HOP, WIN, FUSE>(grid, st, g, tb, bd, wp, a);
  return sc ? launch_one<SRC_NORM_DB, true, HOP, WIN, FUSE>(grid, st, g, tb, bd, wp, a)
            : launch_one<SRC_NORM_DB, false, HOP, WIN, FUSE>(grid, st, g, tb, bd, wp, a);
}

template <int HOP, int WIN>
static const char* launch_geo(int src, bool sc, int grid, cudaStream_t st, const Geo& g, const Tables& tb, const BatchDev& bd,
                              const WpsDev& wp, const FrameArgs& a) {
  return a.wps_iters > 1 ? launch_geo_f<HOP, WIN, true>(src, sc, grid, st, g, tb, bd, wp, a)
                         : launch_geo_f<HOP, WIN, false>(src, sc, grid, st, g, tb, bd, wp, a);
}

bool gl_stream_supported(int hop, int win) {
#define TTSA_X(H, W) if (hop == H && win == W) return WpsGeo<H, W>::kFits;
  TTSA_FIXED_GEOS(TTSA_X)
#undef TTSA_X
  return false;
}

const char* configure_gl_stream() {
  const char* e;
#define TTSA_X(H, W) if ((e = configure_geo<H, W>())) return e;
  TTSA_FIXED_GEOS(TTSA_X)
#undef TTSA_X
  return nullptr;
}

const char* launch_gl_stream(int src, bool sc, int hop, int win, int grid, cudaStream_t st, const Geo& g, const Tables& tb,
                             const BatchDev& bd, const WpsDev& wp, const FrameArgs& a) {
#define TTSA_X(H, W) if (hop == H && win == W) return launch_geo<H, W>(src, sc, grid, st, g, tb, bd, wp, a);
  TTSA_FIXED_GEOS(TTSA_X)
#undef TTSA_X
  return "gl_stream: geometry not instantiated";
}

}  // namespace ttsa
