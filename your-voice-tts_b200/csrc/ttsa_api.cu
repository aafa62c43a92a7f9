// C ABI of libttsa_b200.so (see include/ttsa.h).  Host logic: plan tables, batch layout, kernel orchestration.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <string>
#include <vector>

#include "../../include/ttsa.h"
#include "aux_kernels.cuh"
#include "generic_kernels.cuh"
#include "frame_launch.cuh"
#include "host_tables.hpp"
#include "mel_gemm_tc.cuh"

using namespace ttsa;

// ---------------------------------------------------------------------------------------------------------
// error plumbing
// ---------------------------------------------------------------------------------------------------------
static thread_local std::string g_last_error;
std::atomic<unsigned long long> ttsa::g_launches{0};

static int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
  return code;
}

#define CUDA_TRY(expr)                                                                             \
  do {                                                                                             \
    cudaError_t e_ = (expr);                                                                       \
    if (e_ != cudaSuccess) return fail(TTSA_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(e_));    \
  } while (0)

struct DeviceGuard {
  int prev = -1;
  bool switched = false;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) == cudaSuccess && prev != dev) switched = cudaSetDevice(dev) == cudaSuccess;
  }
  ~DeviceGuard() { if (switched) cudaSetDevice(prev); }
};

// ---------------------------------------------------------------------------------------------------------
// handles
// ---------------------------------------------------------------------------------------------------------
struct ttsa_plan {
  ttsa_config cfg;
  int device = -1;          // -1: host-only plan (tables on the host; no work calls)
  int num_sms = 0;
  Geo geo;
  Tables tb;
  MelParams mel;
  PwParams pw;
  std::vector<double> h_mel;       // [num_mels][F]
  std::vector<double> h_inv_mel;   // [F][num_mels]
  int nz = 32;                     // compile-time zero-pruning class of the frame kernels (20 or 32)
  int ctas_per_sm = 1;
  int debug = 0;                   // profiling builds only (TTSA_PROFILE_BUILD + env TTSA_DEBUG): skip phases of the tile kernel
  int mel_gemm = 0;                // TTSA_MEL_GEMM: 0 default (tensor cores, transposed kernel), 1 "simt", 2 "tc_simple", 3 "tc96" (round-1 pipelined kernel)
  bool generic = false;            // n_fft != 2048: the any-size kernels of generic_kernels.cuh
  GenGeo gg;
  GenTables gt;
  bool fine_ok = true;             // small batches may use the fine-segment form of the tile kernel (TTSA_GL_FINE=0: never)
  bool fixed_geo = true;           // use the kernels compiled for this (hop, win) when they exist (TTSA_GENERIC_GEO=1: never)
  int wps_grid = 0;                // CTAs of the warp-stream kernel (= SMs; TTSA_WPS_GRID overrides it for tests)
  bool feat_stream = false;        // features run the warp-stream kernel (feat_stream.cuh) on batches with enough frames
  bool gl_stream = false;          // Griffin-Lim iterations run the warp-stream kernel (gl_stream.cuh) when the batch has a
                                   // partition for it; TTSA_GL_KERNEL=tile keeps the tile kernel (frame_kernels.cuh)
  // device allocations
  void* d_block = nullptr;         // one allocation holding every table
  const float* d_pinvT = nullptr;  // [num_mels][ldp]
  int ldp = 0;
  // tensor-core operands: B matrices pre-split into 3 bf16 terms in the canonical UMMA layout (mel_gemm_tc.cuh)
  const __nv_bfloat16* d_pinv_tc = nullptr;   // pinv: [5 n-tiles of 208 bins][chunks][3][208 x 80]
  const __nv_bfloat16* d_mel_tc = nullptr;    // mel basis (num_mels == 80 only): [1][13 chunks][3][80 x 80]
  const __nv_bfloat16* d_pinv_tc128 = nullptr; // pinv for the transposed kernel (num_mels <= 80): [9 bin tiles of 128][3][128 x 80]
  const __nv_bfloat16* d_pinv_tc96 = nullptr; // pinv for the pipelined kernel (num_mels <= 80): [11 n-tiles of 96][3][96 x 80]
  int pinv_chunks = 0;
};

struct ttsa_batch {
  int device = -1;
  int B = 0;
  int hop = 0;
  std::vector<int> T, wav_len, tile_off, chunk_off, fine_off;
  std::vector<long long> frame_off, wav_off;
  long long total_frames = 0, total_samples = 0;
  int max_chunks = 0;
  bool dense = true;               // rows of consecutive utterances are contiguous (no per-utterance padding)
  // warp-stream Griffin-Lim partition: one contiguous range of the flattened frame list per warp of a 16-warp CTA per SM
  bool wps_ok = false;
  int wps_grid = 0, wps_win = 0;
  std::vector<int> wps_cut, tsum, wps_u0;
  std::vector<int> tsum_all;       // prefix sum of the frame counts (empty when it overflows int); feat_stream.cuh
  int feat_frames = 0;             // sum of the frame counts (0: not available)
  WpsDev wps_dev{nullptr, nullptr, nullptr};
  void* d_block = nullptr;
  BatchDev dev;
  const int* d_chunk_off = nullptr;
};

// ---------------------------------------------------------------------------------------------------------
// library
// ---------------------------------------------------------------------------------------------------------
extern "C" int ttsa_version(void) { return TTSA_VERSION; }
extern "C" const char* ttsa_last_error(void) { return g_last_error.c_str(); }
extern "C" uint64_t ttsa_launch_count(void) { return g_launches.load(); }

// ---------------------------------------------------------------------------------------------------------
// plan
// ---------------------------------------------------------------------------------------------------------
static int round_up(int x, int m) { return (x + m - 1) / m * m; }

static int kernel_class(const ttsa_config& c) {
  // class 20 ("standard"): <= 20 non-zero rows of packed input and <= 5 window taps per residue mod hop.
  // True for every shipped geometry (275/1102, 200/800, 300/1200); class 32 covers the rest up to win <= 9*hop.
  const int half = (c.win_length + 1) / 2;
  return (half <= 20 * 32 && (c.win_length + c.hop_length - 1) / c.hop_length <= 5) ? 20 : 32;
}

static int build_geo(const ttsa_config& c, Geo& g) {
  const int nz = kernel_class(c);
  std::memset(&g, 0, sizeof(g));
  g.ly = make_layout(c.hop_length, c.win_length, nz);
  g.num_mels = c.num_mels;
  g.preemph = (float)c.preemphasis;

  // value conversions (utils/audio.py:79-126)
  const double LOG2_10 = std::log2(10.0);
  const double mn = c.max_norm, mdb = c.min_level_db, ref = c.ref_level_db;
  double da = 1.0, db0 = 0.0, lo = -INFINITY, hi = INFINITY;
  if (c.signal_norm) {
    if (c.symmetric_norm) { da = -mdb / (2.0 * mn); db0 = mdb / 2.0; if (c.clip_norm) { lo = -mn; hi = mn; } }
    else                  { da = -mdb / mn;         db0 = mdb;       if (c.clip_norm) { lo = 0.0; hi = mn; } }
  }
  const double a_c1 = 0.05 * LOG2_10 * da, a_c0 = 0.05 * LOG2_10 * (db0 + ref);
  g.s_c1 = (float)(c.power * a_c1);
  g.s_c0 = (float)(c.power * a_c0);
  g.s_lo = (float)lo;
  g.s_hi = (float)hi;
  const double cdb = 20.0 * std::log10(2.0);
  double n_a = cdb, n_b = -ref, n_lo = -INFINITY, n_hi = INFINITY;
  if (c.signal_norm) {
    if (c.symmetric_norm) {
      n_a = 2.0 * mn * cdb / -mdb; n_b = 2.0 * mn * (-ref - mdb) / -mdb - mn;
      if (c.clip_norm) { n_lo = -mn; n_hi = mn; }
    } else {
      n_a = mn * cdb / -mdb; n_b = mn * (-ref - mdb) / -mdb;
      if (c.clip_norm) { n_lo = 0.0; n_hi = mn; }
    }
  }
  g.n_a = (float)n_a; g.n_b = (float)n_b; g.n_lo = (float)n_lo; g.n_hi = (float)n_hi;
  g.min_amp = (float)std::pow(10.0, mdb / 20.0);
  return 0;
}

// fp32 -> bf16 round-to-nearest-even (finite inputs), as __float2bfloat16_rn
static uint16_t f2bf(float x) {
  uint32_t u;
  std::memcpy(&u, &x, 4);
  return (uint16_t)((u + 0x7FFFu + ((u >> 16) & 1u)) >> 16);
}
static float bf2f(uint16_t h) {
  const uint32_t u = (uint32_t)h << 16;
  float f;
  std::memcpy(&f, &u, 4);
  return f;
}

// B[n_rows x k_cols] (row-major double) -> [n_tile][chunk][part 3][canonical K-major no-swizzle n_tile x 80] bf16
static std::vector<uint16_t> canon_split_b(const std::vector<double>& B, int n_rows, int k_cols, int n_tile_rows,
                                           int n_tiles, int n_chunks) {
  const size_t part = (size_t)n_tile_rows * kTcChunk;
  std::vector<uint16_t> out((size_t)n_tiles * n_chunks * 3 * part, 0);
  const size_t lbo = (size_t)(n_tile_rows / 8) * 64;     // in bf16 elements (128 bytes per core matrix)
  for (int nt = 0; nt < n_tiles; ++nt)
    for (int ch = 0; ch < n_chunks; ++ch)
      for (int n = 0; n < n_tile_rows; ++n)
        for (int kk = 0; kk < kTcChunk; ++kk) {
          const int row = nt * n_tile_rows + n, k = ch * kTcChunk + kk;
          const float x = (row < n_rows && k < k_cols) ? (float)B[(size_t)row * k_cols + k] : 0.0f;
          const uint16_t h = f2bf(x);
          const float r1 = x - bf2f(h);
          const uint16_t m = f2bf(r1);
          const uint16_t l = f2bf(r1 - bf2f(m));
          const size_t off = (size_t)(kk >> 3) * lbo + (size_t)(n >> 3) * 64 + (size_t)(n & 7) * 8 + (size_t)(kk & 7);
          const size_t base = ((size_t)nt * n_chunks + ch) * 3 * part;
          out[base + off] = h;
          out[base + part + off] = m;
          out[base + 2 * part + off] = l;
        }
  return out;
}

// Tables of the any-size path (generic_kernels.cuh): twiddles, window, banded mel basis, transposed pseudo-inverse.
static int plan_create_generic(ttsa_plan* p, int logn) {
  const ttsa_config& c = p->cfg;
  const int N = c.n_fft, F = c.num_freq;
  GenGeo& g = p->gg;
  g.n_fft = N; g.logn = logn; g.F = F; g.hop = c.hop_length; g.win = c.win_length;
  g.lpad = (N - c.win_length) / 2; g.off0 = N / 2 - g.lpad; g.num_mels = c.num_mels;
  g.preemph = (float)c.preemphasis;
  g.s_c1 = p->geo.s_c1; g.s_c0 = p->geo.s_c0; g.s_lo = p->geo.s_lo; g.s_hi = p->geo.s_hi;
  g.n_a = p->geo.n_a; g.n_b = p->geo.n_b; g.n_lo = p->geo.n_lo; g.n_hi = p->geo.n_hi; g.min_amp = p->geo.min_amp;
  std::vector<float> h_tw((size_t)N), h_win(c.win_length);
  for (int k = 0; k < N / 2; ++k) {
    const double ang = 2.0 * ttsa_host::kPi * k / N;
    h_tw[2 * k] = (float)std::cos(ang);
    h_tw[2 * k + 1] = (float)-std::sin(ang);
  }
  const std::vector<double> w = ttsa_host::hann_periodic(c.win_length);
  for (int m = 0; m < c.win_length; ++m) h_win[m] = (float)w[m];
  std::vector<int> h_lo(c.num_mels, 0), h_cnt(c.num_mels, 0);
  int ld = 1;
  for (int m = 0; m < c.num_mels; ++m) {
    int lo = -1, hi = -1;
    for (int k = 0; k < F; ++k)
      if (p->h_mel[(size_t)m * F + k] != 0.0) { if (lo < 0) lo = k; hi = k; }
    if (lo >= 0) { h_lo[m] = lo; h_cnt[m] = hi - lo + 1; ld = std::max(ld, hi - lo + 1); }
  }
  ld = round_up(ld, 4);
  std::vector<float> h_val((size_t)c.num_mels * ld, 0.f);
  for (int m = 0; m < c.num_mels; ++m)
    for (int cidx = 0; cidx < h_cnt[m]; ++cidx) h_val[(size_t)m * ld + cidx] = (float)p->h_mel[(size_t)m * F + h_lo[m] + cidx];
  p->ldp = round_up(F, 4);
  std::vector<float> h_pinvT((size_t)c.num_mels * p->ldp, 0.f);
  for (int k = 0; k < F; ++k)
    for (int m = 0; m < c.num_mels; ++m) h_pinvT[(size_t)m * p->ldp + k] = (float)p->h_inv_mel[(size_t)k * c.num_mels + m];
  struct Piece { const void* src; size_t bytes; size_t off; };
  std::vector<Piece> pieces = {{h_tw.data(), h_tw.size() * 4, 0}, {h_win.data(), h_win.size() * 4, 0}, {h_lo.data(), h_lo.size() * 4, 0},
                               {h_cnt.data(), h_cnt.size() * 4, 0}, {h_val.data(), h_val.size() * 4, 0},
                               {h_pinvT.data(), h_pinvT.size() * 4, 0}};
  size_t total = 0;
  for (auto& pc : pieces) { pc.off = total; total += (pc.bytes + 255) / 256 * 256; }
  if (cudaMalloc(&p->d_block, total) != cudaSuccess) return fail(TTSA_ERR_CUDA, "cudaMalloc(%zu) for plan tables failed", total);
  for (auto& pc : pieces) {
    cudaError_t e = cudaMemcpy((char*)p->d_block + pc.off, pc.src, pc.bytes, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { cudaFree(p->d_block); p->d_block = nullptr; return fail(TTSA_ERR_CUDA, "table upload: %s", cudaGetErrorString(e)); }
  }
  char* base = (char*)p->d_block;
  p->gt.tw = (const float2*)(base + pieces[0].off);
  p->gt.win = (const float*)(base + pieces[1].off);
  p->gt.mel_lo = (const int*)(base + pieces[2].off);
  p->gt.mel_cnt = (const int*)(base + pieces[3].off);
  p->gt.mel_val = (const float*)(base + pieces[4].off);
  p->gt.mel_ld = ld;
  p->d_pinvT = (const float*)(base + pieces[5].off);
  p->ctas_per_sm = 1;
  cudaError_t e = cudaFuncSetAttribute(mel_to_linear_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       kMtlMaxK * (kMtlBins + kMtlRows) * 4);
  if (e != cudaSuccess) { cudaFree(p->d_block); p->d_block = nullptr; return fail(TTSA_ERR_CUDA, "mel kernel configuration: %s", cudaGetErrorString(e)); }
  return TTSA_OK;
}

// launch one any-size frame kernel
static int gen_launch(const ttsa_plan* plan, const ttsa_batch* batch, int mode, int src, FrameArgs a, cudaStream_t st) {
  if (batch->total_frames == 0) return TTSA_OK;
  a.rows_total = batch->total_frames;
  const int grid = (int)std::min<long long>(batch->total_frames, (long long)plan->num_sms * 8);
  const size_t smem = (size_t)plan->gg.n_fft * 8 + (size_t)(plan->gg.F + 1) * 4;
  const GenGeo& g = plan->gg; const GenTables& t = plan->gt; const BatchDev& bd = batch->dev;
#define TTSA_GEN(M, S) gen_frame_kernel<M, S><<<grid, kGenThreads, smem, st>>>(g, t, bd, a)
  if (mode == MODE_ANALYSIS) { if (src == OUT_COMPLEX) TTSA_GEN(MODE_ANALYSIS, OUT_COMPLEX); else TTSA_GEN(MODE_ANALYSIS, OUT_FEATURES); }
  else if (mode == MODE_GL_ITER) { if (src == SRC_MAG) TTSA_GEN(MODE_GL_ITER, SRC_MAG); else TTSA_GEN(MODE_GL_ITER, SRC_NORM_DB); }
  else { if (src == SRC_MAG) TTSA_GEN(MODE_SYNTH, SRC_MAG); else if (src == SRC_NORM_DB) TTSA_GEN(MODE_SYNTH, SRC_NORM_DB); else TTSA_GEN(MODE_SYNTH, SRC_COMPLEX); }
#undef TTSA_GEN
  g_launches += 1;
  CUDA_TRY(cudaGetLastError());
  return TTSA_OK;
}

// zero the waveform, overlap-add every frame into it, divide by the window sum of squares
static int gen_synthesise(const ttsa_plan* plan, const ttsa_batch* batch, int mode, int src, const FrameArgs& a, cudaStream_t st) {
  if (batch->total_samples == 0) return TTSA_OK;
  CUDA_TRY(cudaMemsetAsync(a.wav_out, 0, (size_t)batch->total_samples * 4, st));
  // overlap-add without atomics: frames t = p (mod R), R = ceil(win / hop), do not overlap each other, so launch phase p
  // adds them with plain read-modify-write; the phases run in stream order -> one fixed summation order per sample
  const int phases = (plan->cfg.win_length + plan->cfg.hop_length - 1) / plan->cfg.hop_length;
  for (int ph = 0; ph < phases; ++ph) {
    FrameArgs b = a;
    b.ola_phase = ph; b.ola_phases = phases;
    if (int rc = gen_launch(plan, batch, mode, src, b, st)) return rc;
  }
  int maxlen = 0;
  for (int v : batch->wav_len) maxlen = std::max(maxlen, v);
  if (maxlen == 0) return TTSA_OK;
  dim3 grid((unsigned)std::max(1, std::min(1024, (maxlen + 255) / 256)), (unsigned)batch->B);
  gen_wss_kernel<<<grid, 256, 0, st>>>(plan->gg, plan->gt, batch->dev, a.wav_out);
  g_launches += 1;
  CUDA_TRY(cudaGetLastError());
  return TTSA_OK;
}

extern "C" int ttsa_plan_create(const ttsa_config* cfg, int device, ttsa_plan** out) {
  if (!cfg || !out) return fail(TTSA_ERR_BAD_ARG, "null argument");
  *out = nullptr;
  const ttsa_config& c = *cfg;
  if (c.sample_rate <= 0 || c.num_mels <= 0 || c.num_freq < 2)
    return fail(TTSA_ERR_BAD_CONFIG, "sample_rate, num_mels and num_freq must be positive");
  if (c.n_fft != (c.num_freq - 1) * 2) return fail(TTSA_ERR_BAD_CONFIG, "n_fft must equal (num_freq - 1) * 2");
  if (c.mel_fmax > 0 && c.mel_fmax > c.sample_rate / 2)   // assert at utils/audio.py:70-71 (integer sr // 2)
    return fail(TTSA_ERR_BAD_CONFIG, "mel_fmax %.1f > sample_rate // 2 = %d", c.mel_fmax, c.sample_rate / 2);
  const bool generic = c.n_fft != kNfft;
  int logn = 0;
  while ((1 << logn) < c.n_fft) ++logn;
  if (generic && ((1 << logn) != c.n_fft || c.n_fft < 256 || c.n_fft > 4096))
    return fail(TTSA_ERR_UNSUPPORTED, "num_freq %d: n_fft = %d must be a power of two in [256, 4096]", c.num_freq, c.n_fft);
  if (c.hop_length < 2 || c.win_length < c.hop_length || c.win_length > c.n_fft)
    return fail(TTSA_ERR_UNSUPPORTED, "need 2 <= hop_length <= win_length <= n_fft (hop %d, win %d)", c.hop_length, c.win_length);
  if (!generic && c.win_length - c.hop_length > kNF * c.hop_length)
    return fail(TTSA_ERR_UNSUPPORTED, "win_length %d > %d * hop_length %d", c.win_length, kNF + 1, c.hop_length);
  if (c.num_mels > kMtlMaxK) return fail(TTSA_ERR_UNSUPPORTED, "num_mels %d > %d", c.num_mels, kMtlMaxK);
  if (c.signal_norm && (c.max_norm <= 0 || c.min_level_db >= 0))
    return fail(TTSA_ERR_BAD_CONFIG, "signal_norm needs max_norm > 0 and min_level_db < 0");

  ttsa_plan* p = new ttsa_plan();
  p->cfg = c;
  p->device = device;
  p->generic = generic;
  build_geo(c, p->geo);
  p->nz = kernel_class(c);
  if (!generic && (size_t)p->geo.ly.sm_total * 4 > 227 * 1024) {
    delete p;
    return fail(TTSA_ERR_UNSUPPORTED, "hop/win need %d bytes of shared memory per CTA", p->geo.ly.sm_total * 4);
  }

  // ---- host tables (float64) ----
  const double fmax = c.mel_fmax > 0 ? c.mel_fmax : 0.5 * c.sample_rate;
  p->h_mel = ttsa_host::mel_basis(c.sample_rate, c.n_fft, c.num_mels, c.mel_fmin, fmax);
  p->h_inv_mel = ttsa_host::pinv_wide(p->h_mel, c.num_mels, c.num_freq);

  // elementwise parameter blocks
  p->pw.signal_norm = c.signal_norm; p->pw.symmetric_norm = c.symmetric_norm; p->pw.clip_norm = c.clip_norm;
  p->pw.min_level_db = (float)c.min_level_db; p->pw.max_norm = (float)c.max_norm; p->pw.min_amp = p->geo.min_amp;
  p->pw.op = 0;
  MelParams& mp = p->mel;
  mp.a_c1 = (float)(p->geo.s_c1 / c.power); mp.a_c0 = (float)(p->geo.s_c0 / c.power);
  {  // recompute without the power factor to avoid the division rounding (and power == 0)
    Geo g1; ttsa_config c1 = c; c1.power = 1.0; build_geo(c1, g1);
    mp.a_c1 = g1.s_c1; mp.a_c0 = g1.s_c0;
  }
  mp.a_lo = p->geo.s_lo; mp.a_hi = p->geo.s_hi;
  mp.n_a = p->geo.n_a; mp.n_b = p->geo.n_b; mp.n_lo = p->geo.n_lo; mp.n_hi = p->geo.n_hi; mp.min_amp = p->geo.min_amp;
  mp.power = (float)c.power; mp.num_mels = c.num_mels; mp.F = c.num_freq; mp.rows = 0;

  if (device < 0) { *out = p; return TTSA_OK; }

  // ---- device tables ----
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device >= ndev) {
    delete p;
    return fail(TTSA_ERR_NO_DEVICE, "CUDA device %d not available (%d devices)", device, ndev);
  }
  DeviceGuard guard(device);
  cudaDeviceProp prop;
  {
    const cudaError_t pe = cudaGetDeviceProperties(&prop, device);
    if (pe != cudaSuccess) {
      delete p;
      return fail(TTSA_ERR_CUDA, "cudaGetDeviceProperties(%d): %s", device, cudaGetErrorString(pe));
    }
  }
  if (prop.major != 10) {
    delete p;
    return fail(TTSA_ERR_NO_DEVICE, "device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
  }
  p->num_sms = prop.multiProcessorCount;
  if (generic) {
    if (int rc = plan_create_generic(p, logn)) { delete p; return rc; }
    *out = p;
    return TTSA_OK;
  }

  const std::vector<double> w = ttsa_host::hann_periodic(c.win_length);
  std::vector<float> h_tw(2048), h_g(1024), h_wE(1024, 0.f), h_wO(1024, 0.f), h_wE2(1024, 0.f), h_wO2(1024, 0.f),
      h_pw(round_up(c.hop_length, 4), 1.f);
  for (int m = 0; m < 16; ++m)            // tw4[m][lane] = (wr(2m), wr(2m+1), wi(2m), wi(2m+1)), w(k) = W_1024^(lane k)
    for (int lane = 0; lane < 32; ++lane)
      for (int e = 0; e < 2; ++e) {
        const double ang = 2.0 * ttsa_host::kPi * ((lane * (2 * m + e)) % 1024) / 1024.0;
        h_tw[4 * (m * 32 + lane) + e] = (float)std::cos(ang);
        h_tw[4 * (m * 32 + lane) + 2 + e] = (float)-std::sin(ang);
      }
  for (int m = 0; m < 8; ++m)             // g4[m][lane] = (gx(2m), gx(2m+1), gy(2m), gy(2m+1)), g(k1) = -j W_2048^(32 k1 + lane)
    for (int lane = 0; lane < 32; ++lane)
      for (int e = 0; e < 2; ++e) {
        const double ang = ttsa_host::kPi * (32 * (2 * m + e) + lane) / 1024.0;
        h_g[4 * (m * 32 + lane) + e] = (float)-std::sin(ang);
        h_g[4 * (m * 32 + lane) + 2 + e] = (float)-std::cos(ang);
      }
  for (int m = 0; m < c.win_length; ++m) ((m & 1) ? h_wO : h_wE)[m >> 1] = (float)w[m];
  for (int m = 0; m < 16; ++m)            // paired taps for two packed rows: q = lane + 64 m and q + 32
    for (int lane = 0; lane < 32; ++lane) {
      const int q = lane + 64 * m;
      h_wE2[2 * (m * 32 + lane)] = h_wE[q];
      h_wE2[2 * (m * 32 + lane) + 1] = q + 32 < 1024 ? h_wE[q + 32] : 0.f;
      h_wO2[2 * (m * 32 + lane)] = h_wO[q];
      h_wO2[2 * (m * 32 + lane) + 1] = q + 32 < 1024 ? h_wO[q + 32] : 0.f;
    }
  for (int r = 0; r < c.hop_length; ++r) {
    double acc = 0.0;
    for (int m = r; m < c.win_length; m += c.hop_length) { const double wf = (double)(float)w[m]; acc += wf * wf; }
    h_pw[r] = (float)(1.0 / acc);
  }
  std::vector<float> h_pw2(2 * (size_t)c.hop_length);          // sample-pair form with the 1/n_fft of the inverse FFT (gl_stream.cuh)
  for (int r = 0; r < c.hop_length; ++r) {
    h_pw2[2 * r] = h_pw[r] * (1.0f / (float)kNfft);
    h_pw2[2 * r + 1] = h_pw[(r + 1) % c.hop_length] * (1.0f / (float)kNfft);
  }
  // 1 / (n_fft * window sum of squares) at the edges of an utterance (gl_stream.cuh): sample i < warm*hop - win/2 lacks the
  // frames before frame 0, sample L - ntail + j lacks frame T (the frames are at t*hop - win/2, L = hop*(T-1))
  std::vector<float> h_edge_head, h_edge_tail;
  {
    const int hop = c.hop_length, win = c.win_length, warm = (win - 1) / hop;
    const int nhead = std::max(0, warm * hop - win / 2), ntail = std::max(0, win / 2 - hop);
    h_edge_head.assign(std::max(nhead, 1), 0.f);
    h_edge_tail.assign(std::max(ntail, 1), 0.f);
    // the kernel adds the terms from the latest frame down (same order here: bit-identical to its on-the-fly sum)
    auto inv_wss_desc = [&](long long i, long long t_lo, long long t_hi) {   // frames t_lo..t_hi exist
      float ws = 0.f;
      for (long long tt = t_hi; tt >= t_lo; --tt) {
        const long long m = i - (tt * hop - win / 2);
        if (m >= 0 && m < win) { const float wf = (float)w[m]; ws = std::fmaf(wf, wf, ws); }
      }
      return ws > 1.17549435e-38f ? (1.0f / (float)kNfft) / ws : 1.0f / (float)kNfft;
    };
    const long long Tbig = 1000;                                        // any utterance long enough for both tables
    for (int i = 0; i < nhead; ++i) h_edge_head[i] = inv_wss_desc(i, 0, Tbig - 1);
    const long long Lbig = (long long)hop * (Tbig - 1);
    for (int j = 0; j < ntail; ++j) h_edge_tail[j] = inv_wss_desc(Lbig - ntail + j, 0, Tbig - 1);
  }
  // table image of the warp-stream Griffin-Lim kernel (gl_stream.cuh, WpsGeo<hop, win>): tw4 | g4 | wE | wO1 | pwx
  std::vector<float> h_wps;
  {
    const int np = c.win_length / 2 + (c.hop_length & 1), rh = (np + 31) / 32 * 32;
    const int emit_rows = ((c.hop_length + 1) / 2 + 31) / 32, npwx = (64 * emit_rows + 4 + 3) / 4 * 4;
    // window taps per LANE (kWpsWinStride consecutive floats: pair q = lane + 32 n at [lane][n], so a lane fetches its taps
    // with 16-byte loads): wA[lane][n] = w[2q]; wB[r][n] = w[2 (r - 1 + 32 n) + 1], rows r = 0..32, i.e. row lane + 1 holds
    // the odd taps of pair q and row lane those of pair q - 1 (frames that start one sample early); w = 0 outside the window
    (void)rh;
    h_wps.assign(2048 + 1024 + 65 * (size_t)kWpsWinStride + npwx, 0.f);
    float* q = h_wps.data();
    for (int i = 0; i < 2048; ++i) q[i] = h_tw[i];
    q += 2048;
    for (int i = 0; i < 1024; ++i) q[i] = h_g[i];
    q += 1024;
    for (int lane = 0; lane < 32; ++lane)
      for (int n = 0; n < kWpsWinStride; ++n) q[lane * kWpsWinStride + n] = lane + 32 * n < 1024 ? h_wE[lane + 32 * n] : 0.f;
    q += 32 * kWpsWinStride;
    for (int r = 0; r < 33; ++r)
      for (int n = 0; n < kWpsWinStride; ++n) {
        const int pq = r - 1 + 32 * n;
        q[r * kWpsWinStride + n] = (pq >= 0 && pq < 1024) ? h_wO[pq] : 0.f;
      }
    q += 33 * kWpsWinStride;
    for (int j = 0; j < npwx; ++j) q[j] = h_pw[(j - 1 + c.hop_length) % c.hop_length] * (1.0f / (float)kNfft);
  }
  // shared-memory image of the frame kernels' constant tables (Layout: [sm_wE, sm_mbar))
  const Layout& ly = p->geo.ly;
  std::vector<float> h_img(ly.image_floats, 0.f);
  {
    float* img = h_img.data() - ly.sm_wE;
    for (int i = 0; i < ly.wlen; ++i) { img[ly.sm_wE + i] = h_wE2[i]; img[ly.sm_wO + i] = h_wO2[i]; }
    for (int r = 0; r < c.hop_length; ++r) img[ly.sm_pw + r] = h_pw[r] * (1.0f / (float)kNfft);   // 1/wss and the 1/n_fft of the inverse FFT
    for (int m = 0; m < c.win_length; ++m) img[ly.sm_wsyn + m] = (m & 1) ? -h_wO[m >> 1] : h_wE[m >> 1];
    for (int i = 0; i < 2048; ++i) img[ly.sm_tw + i] = h_tw[i];
    for (int i = 0; i < 1024; ++i) img[ly.sm_g + i] = h_g[i];
  }
  // banded mel basis
  std::vector<int> h_lo(c.num_mels, 0), h_cnt(c.num_mels, 0);
  int ld = 1;
  for (int m = 0; m < c.num_mels; ++m) {
    int lo = -1, hi = -1;
    for (int k = 0; k < kF; ++k)
      if (p->h_mel[(size_t)m * kF + k] != 0.0) { if (lo < 0) lo = k; hi = k; }
    if (lo >= 0) { h_lo[m] = lo; h_cnt[m] = hi - lo + 1; ld = std::max(ld, hi - lo + 1); }
  }
  ld = round_up(ld, 4);
  std::vector<float> h_val((size_t)c.num_mels * ld, 0.f);
  for (int m = 0; m < c.num_mels; ++m)
    for (int cidx = 0; cidx < h_cnt[m]; ++cidx) h_val[(size_t)m * ld + cidx] = (float)p->h_mel[(size_t)m * kF + h_lo[m] + cidx];
  p->ldp = round_up(kF, 4);
  std::vector<float> h_pinvT((size_t)c.num_mels * p->ldp, 0.f);
  for (int k = 0; k < kF; ++k)
    for (int m = 0; m < c.num_mels; ++m) h_pinvT[(size_t)m * p->ldp + k] = (float)p->h_inv_mel[(size_t)k * c.num_mels + m];

  // compact copy of the banded basis for the feature kernel's shared memory: int2 (first tap, first bin) per filter
  // (+ one end entry), then the taps back to back; used when it fits beside the frame buffers with 2 CTAs per SM
  std::vector<float> h_melc;
  {
    std::vector<int> desc(2 * (c.num_mels + 1), 0);
    std::vector<float> taps;
    for (int m = 0; m < c.num_mels; ++m) {
      desc[2 * m] = (int)taps.size(); desc[2 * m + 1] = h_lo[m];
      for (int cidx = 0; cidx < h_cnt[m]; ++cidx) taps.push_back(h_val[(size_t)m * ld + cidx]);
    }
    desc[2 * c.num_mels] = (int)taps.size();
    h_melc.resize(desc.size() + taps.size());
    std::memcpy(h_melc.data(), desc.data(), desc.size() * 4);
    std::memcpy(h_melc.data() + desc.size(), taps.data(), taps.size() * 4);
    const size_t per_cta_limit = (233472 - 2 * 1024) / 2;          // two CTAs per SM, 1 KB reserved each
    const bool fits = ((size_t)ly.sm_total + h_melc.size()) * 4 <= per_cta_limit;
    p->geo.mel_smem_floats = fits ? (int)h_melc.size() : 0;
  }
  // Lane schedule of the banded basis for the warp-stream feature kernel (feat_stream.cuh): every lane owns up to three
  // filters (longest-first onto the least loaded lane, so all lanes walk about sum(taps) / 32 taps instead of the three
  // longest filters back to back) and walks its taps in an order chosen so that the 32 lanes of a step read 32 different
  // shared-memory banks of the magnitude row.  Entry [step][lane] = (w0, w1, w2, bin): the tap's weight sits in the slot of
  // the lane's filter it belongs to, the other two are zero, so a step is one 16-byte load, one magnitude load and three
  // FMAs with no branch; then [3][32] filter indices (-1: none).
  std::vector<float> h_msched;
  p->geo.mel_steps = 0;
  if (c.num_mels <= 96) {
    struct Tap { int bin; float w; int slot; };
    std::vector<std::vector<Tap>> lane_taps(32);
    std::vector<int> lane_nf(32, 0), fid(96, -1);
    std::vector<int> order(c.num_mels);
    for (int m = 0; m < c.num_mels; ++m) order[m] = m;
    std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return h_cnt[x] > h_cnt[y]; });
    for (int m : order) {
      int best = -1;
      for (int l = 0; l < 32; ++l)
        if (lane_nf[l] < 3 && (best < 0 || lane_taps[l].size() < lane_taps[best].size())) best = l;
      const int slot = lane_nf[best]++;
      fid[slot * 32 + best] = m;
      for (int cidx = 0; cidx < h_cnt[m]; ++cidx) {
        const float w = h_val[(size_t)m * ld + cidx];
        if (w != 0.0f) lane_taps[best].push_back({h_lo[m] + cidx, w, slot});
      }
    }
    size_t steps = 1;
    for (int l = 0; l < 32; ++l) steps = std::max(steps, lane_taps[l].size());
    h_msched.assign(steps * 32 * 4 + 96, 0.f);
    std::vector<std::vector<char>> done(32);
    for (int l = 0; l < 32; ++l) done[l].assign(lane_taps[l].size(), 0);
    std::vector<size_t> left(32);
    for (int l = 0; l < 32; ++l) left[l] = lane_taps[l].size();
    for (size_t st = 0; st < steps; ++st) {
      unsigned used = 0;                                            // banks read in this step
      std::vector<int> lanes(32);
      for (int l = 0; l < 32; ++l) lanes[l] = l;
      // lanes that cannot afford an idle step choose first
      std::stable_sort(lanes.begin(), lanes.end(), [&](int x, int y) { return left[x] > left[y]; });
      for (int l : lanes) {
        float* e = &h_msched[(st * 32 + l) * 4];
        const bool must = left[l] >= steps - st;                    // no idle step left for this lane
        int pick = -1;
        for (size_t i = 0; i < lane_taps[l].size(); ++i)
          if (!done[l][i] && !(used >> (lane_taps[l][i].bin & 31) & 1u)) { pick = (int)i; break; }
        if (pick < 0 && must)
          for (size_t i = 0; i < lane_taps[l].size(); ++i) if (!done[l][i]) { pick = (int)i; break; }   // a bank conflict
        int bin = 0;
        if (pick >= 0) {
          const Tap& t = lane_taps[l][pick];
          done[l][pick] = 1; --left[l];
          e[t.slot] = t.w;
          bin = t.bin;
        } else {                                                    // idle step: zero weights, a bin in a free bank
          for (int b = 0; b < 32; ++b) if (!(used >> b & 1u)) { bin = b; break; }
        }
        used |= 1u << (bin & 31);
        std::memcpy(&e[3], &bin, 4);
      }
    }
    bool all = true;
    for (int l = 0; l < 32; ++l) all = all && left[l] == 0;
    if (all) {
      std::memcpy(&h_msched[steps * 32 * 4], fid.data(), 96 * 4);
      p->geo.mel_steps = (int)steps;

    } else {
      h_msched.clear();
    }
  }
  // segment schedule of the same basis (host_tables.hpp): one read of every bin, two accumulators per cell; the default of the
  // warp-stream feature kernel when the basis is a 50 %-overlap triangular bank (TTSA_FEAT_MEL=lane keeps the lane schedule)
  std::vector<uint32_t> h_mseg;
  p->geo.mel_seg_pairs[0] = p->geo.mel_seg_pairs[1] = p->geo.mel_seg_pairs[2] = 0;
  {
    const char* fm = std::getenv("TTSA_FEAT_MEL");
    if (!(fm != nullptr && std::strcmp(fm, "lane") == 0)) {
      ttsa_host::MelSegSchedule sc = ttsa_host::mel_segment_schedule(p->h_mel, c.num_mels, kF);
      if (sc.ok) {
        h_mseg.swap(sc.words);
        for (int i = 0; i < 3; ++i) p->geo.mel_seg_pairs[i] = sc.pairs[i];
      }
    }
  }
  p->pinv_chunks = (c.num_mels + kTcChunk - 1) / kTcChunk;
  const std::vector<uint16_t> h_pinv_tc = canon_split_b(p->h_inv_mel, kF, c.num_mels, 208, 5, p->pinv_chunks);
  std::vector<uint16_t> h_mel_tc;
  if (c.num_mels == 80) h_mel_tc = canon_split_b(p->h_mel, 80, kF, 80, 1, 13);
  std::vector<uint16_t> h_pinv_tc96;
  if (p->pinv_chunks == 1) h_pinv_tc96 = canon_split_b(p->h_inv_mel, kF, c.num_mels, kM2lN, kM2lTiles, 1);
  std::vector<uint16_t> h_pinv_tc128;
  if (p->pinv_chunks == 1) h_pinv_tc128 = canon_split_b(p->h_inv_mel, kF, c.num_mels, kT2Bins, kT2Tiles, 1);
  // one device block
  struct Piece { const void* src; size_t bytes; size_t off; };
  std::vector<Piece> pieces = {
      {h_tw.data(), h_tw.size() * 4, 0}, {h_g.data(), h_g.size() * 4, 0}, {h_wE.data(), h_wE.size() * 4, 0},
      {h_wO.data(), h_wO.size() * 4, 0}, {h_pw.data(), h_pw.size() * 4, 0}, {h_lo.data(), h_lo.size() * 4, 0},
      {h_cnt.data(), h_cnt.size() * 4, 0}, {h_val.data(), h_val.size() * 4, 0}, {h_pinvT.data(), h_pinvT.size() * 4, 0},
      {h_wE2.data(), h_wE2.size() * 4, 0}, {h_wO2.data(), h_wO2.size() * 4, 0},
      {h_pinv_tc.data(), h_pinv_tc.size() * 2, 0}, {h_mel_tc.data(), h_mel_tc.size() * 2, 0},
      {h_pinv_tc96.data(), h_pinv_tc96.size() * 2, 0}, {h_img.data(), h_img.size() * 4, 0},
      {h_melc.data(), h_melc.size() * 4, 0}, {h_pw2.data(), h_pw2.size() * 4, 0}, {h_wps.data(), h_wps.size() * 4, 0},
      {h_edge_head.data(), h_edge_head.size() * 4, 0}, {h_edge_tail.data(), h_edge_tail.size() * 4, 0},
      {h_msched.data(), h_msched.size() * 4, 0}, {h_pinv_tc128.data(), h_pinv_tc128.size() * 2, 0},
      {h_mseg.data(), h_mseg.size() * 4, 0}};
  size_t total = 0;
  for (auto& pc : pieces) { pc.off = total; total += (pc.bytes + 255) / 256 * 256; }
  if (cudaMalloc(&p->d_block, total) != cudaSuccess) {
    delete p;
    return fail(TTSA_ERR_CUDA, "cudaMalloc(%zu) for plan tables failed", total);
  }
  for (auto& pc : pieces) {
    cudaError_t e = cudaMemcpy((char*)p->d_block + pc.off, pc.src, pc.bytes, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { cudaFree(p->d_block); delete p; return fail(TTSA_ERR_CUDA, "table upload: %s", cudaGetErrorString(e)); }
  }
  char* base = (char*)p->d_block;
  p->tb.tw4 = (const float4*)(base + pieces[0].off);
  p->tb.g4 = (const float4*)(base + pieces[1].off);
  p->tb.wE2 = (const float2*)(base + pieces[9].off);
  p->tb.wO2 = (const float2*)(base + pieces[10].off);
  p->d_pinv_tc = (const __nv_bfloat16*)(base + pieces[11].off);
  p->d_mel_tc = h_mel_tc.empty() ? nullptr : (const __nv_bfloat16*)(base + pieces[12].off);
  p->d_pinv_tc96 = h_pinv_tc96.empty() ? nullptr : (const __nv_bfloat16*)(base + pieces[13].off);
  p->d_pinv_tc128 = h_pinv_tc128.empty() ? nullptr : (const __nv_bfloat16*)(base + pieces[21].off);
  p->tb.smem_image = (const float*)(base + pieces[14].off);
  p->tb.mel_compact = (const float*)(base + pieces[15].off);
  p->tb.wE = (const float*)(base + pieces[2].off);
  p->tb.wO = (const float*)(base + pieces[3].off);
  p->tb.pw = (const float*)(base + pieces[4].off);
  p->tb.pw2 = (const float2*)(base + pieces[16].off);
  p->tb.wps_image = (const float*)(base + pieces[17].off);
  p->tb.mel_sched = h_msched.empty() ? nullptr : (const float*)(base + pieces[20].off);
  p->tb.mel_seg = h_mseg.empty() ? nullptr : (const unsigned*)(base + pieces[22].off);
  p->tb.edge_head = (const float*)(base + pieces[18].off);
  p->tb.edge_tail = (const float*)(base + pieces[19].off);
  p->tb.mel_lo = (const int*)(base + pieces[5].off);
  p->tb.mel_cnt = (const int*)(base + pieces[6].off);
  p->tb.mel_val = (const float*)(base + pieces[7].off);
  p->tb.mel_ld = ld;
  p->d_pinvT = (const float*)(base + pieces[8].off);

  // configure every kernel once (dynamic shared memory opt-in, occupancy), outside any stream capture
  const size_t smem_bytes = ((size_t)p->geo.ly.sm_total + p->geo.mel_smem_floats) * 4;   // the feature kernel's request
  int occ = 0;
  const char* err = configure_frame_kernels(smem_bytes, &occ);
  if (err) { cudaFree(p->d_block); delete p; return fail(TTSA_ERR_CUDA, "kernel configuration: %s", err); }
  p->ctas_per_sm = occ < 1 ? 1 : occ;
  if (gl_stream_supported(c.hop_length, c.win_length)) {
    const char* gk = std::getenv("TTSA_GL_KERNEL");
    p->gl_stream = !(gk != nullptr && std::strcmp(gk, "tile") == 0);
    const char* wg = std::getenv("TTSA_WPS_GRID");
    p->wps_grid = (wg != nullptr && std::atoi(wg) > 0) ? std::min(std::atoi(wg), p->num_sms) : p->num_sms;
    if (p->gl_stream) {
      err = configure_gl_stream();
      if (err) { cudaFree(p->d_block); delete p; return fail(TTSA_ERR_CUDA, "kernel configuration: %s", err); }
    }
  }
  if (feat_stream_supported(c.hop_length, c.win_length, feat_mel_floats(p->geo))) {
    const char* fk = std::getenv("TTSA_FEAT_KERNEL");
    p->feat_stream = !(fk != nullptr && std::strcmp(fk, "tile") == 0);
    const char* wg = std::getenv("TTSA_WPS_GRID");
    p->wps_grid = (wg != nullptr && std::atoi(wg) > 0) ? std::min(std::atoi(wg), p->num_sms) : p->num_sms;
    if (p->feat_stream) {
      err = configure_feat_stream();
      if (err) { cudaFree(p->d_block); delete p; return fail(TTSA_ERR_CUDA, "kernel configuration: %s", err); }
    }
  }
  { const char* gen = std::getenv("TTSA_GENERIC_GEO"); p->fixed_geo = !(gen != nullptr && std::atoi(gen) != 0); }
  { const char* fi = std::getenv("TTSA_GL_FINE"); p->fine_ok = !(fi != nullptr && std::atoi(fi) == 0); }
#ifdef TTSA_PROFILE_BUILD
  { const char* dbg = std::getenv("TTSA_DEBUG"); p->debug = dbg ? std::atoi(dbg) : 0; }
#endif
  { const char* mg = std::getenv("TTSA_MEL_GEMM");
    p->mel_gemm = mg == nullptr ? 0 : (std::strcmp(mg, "simt") == 0 ? 1 : (std::strcmp(mg, "tc_simple") == 0 ? 2 : (std::strcmp(mg, "tc96") == 0 ? 3 : 0))); }
  cudaError_t e = cudaFuncSetAttribute(mel_to_linear_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       kMtlMaxK * (kMtlBins + kMtlRows) * 4);
  if (e != cudaSuccess) { cudaFree(p->d_block); delete p; return fail(TTSA_ERR_CUDA, "mel kernel configuration: %s", cudaGetErrorString(e)); }
  e = cudaFuncSetAttribute(gemm_bf16x3_tc_kernel<208>, cudaFuncAttributeMaxDynamicSharedMemorySize, 3 * kTcRows * kTcChunk * 2 + 3 * 208 * kTcChunk * 2);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(gemm_bf16x3_tc_kernel<80>, cudaFuncAttributeMaxDynamicSharedMemorySize, 3 * kTcRows * kTcChunk * 2 + 3 * 80 * kTcChunk * 2);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(mel_to_linear_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kM2lSmem);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(mel_to_linear_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kT2Smem);
  if (e != cudaSuccess) { cudaFree(p->d_block); delete p; return fail(TTSA_ERR_CUDA, "tensor-core GEMM configuration: %s", cudaGetErrorString(e)); }
  *out = p;
  return TTSA_OK;
}

extern "C" int ttsa_plan_destroy(ttsa_plan* plan) {
  if (!plan) return TTSA_OK;
  if (plan->d_block) { DeviceGuard g(plan->device); cudaFree(plan->d_block); }
  delete plan;
  return TTSA_OK;
}

extern "C" int ttsa_plan_mel_basis(const ttsa_plan* plan, double* host_out) {
  if (!plan || !host_out) return fail(TTSA_ERR_BAD_ARG, "null argument");
  std::memcpy(host_out, plan->h_mel.data(), plan->h_mel.size() * sizeof(double));
  return TTSA_OK;
}

extern "C" int ttsa_plan_inv_mel_basis(const ttsa_plan* plan, double* host_out) {
  if (!plan || !host_out) return fail(TTSA_ERR_BAD_ARG, "null argument");
  std::memcpy(host_out, plan->h_inv_mel.data(), plan->h_inv_mel.size() * sizeof(double));
  return TTSA_OK;
}

extern "C" int64_t ttsa_plan_mel_schedule(const ttsa_plan* plan, int32_t* pairs_out, uint32_t* words_out, int64_t cap_words) {
  if (!plan || !pairs_out) return fail(TTSA_ERR_BAD_ARG, "null argument");
  const ttsa_host::MelSegSchedule sc = ttsa_host::mel_segment_schedule(plan->h_mel, plan->cfg.num_mels, plan->cfg.num_freq);
  for (int i = 0; i < 3; ++i) pairs_out[i] = sc.ok ? sc.pairs[i] : 0;
  if (!sc.ok) return 0;
  if (words_out != nullptr) {
    if ((int64_t)sc.words.size() > cap_words) return fail(TTSA_ERR_BAD_ARG, "schedule needs %zu words, buffer holds %lld", sc.words.size(), (long long)cap_words);
    std::memcpy(words_out, sc.words.data(), sc.words.size() * 4);
  }
  return (int64_t)sc.words.size();
}

// Work partition of the warp-stream Griffin-Lim kernel: the flattened frame list (utterance after utterance) is cut into
// one contiguous range per warp.  A range that starts inside an utterance hands its first win - hop samples over to the
// owner of the frames before the cut (gl_stream.cuh), which works when that owner holds ALL `warm` frames that overlap
// the cut: a run that both starts and ends inside one utterance must be at least `minrun` frames long, and no cut may lie
// closer than `minrun` frames to an utterance boundary.  Balance matters more than anything else here (the launch ends
// with its longest run): cuts start evenly spaced, cuts too close to a boundary snap onto it, and the cuts between two
// consecutive boundary cuts of the same utterance are then re-spaced evenly, so run lengths differ by at most one frame
// wherever utterances are longer than a run.  Returns false when the batch is too small (the tile kernel serves it).
// extra cost of an utterance's first (which = 0) / last (1) run in frame times; TTSA_WPS_EDGE="e0,e1" overrides (experiments)
static double wps_edge_cost(int which) {
  double e[2] = {0.0, 0.0};   // measured at 64 x 482: (0, 0) 0.1280 ms, (1, 1.5) 0.1287 ms, (2, 2.5) 0.1296 ms per launch
  if (const char* s = std::getenv("TTSA_WPS_EDGE")) {
    double x0, x1;
    if (std::sscanf(s, "%lf,%lf", &x0, &x1) == 2 && x0 >= 0 && x1 >= 0 && x0 < 8 && x1 < 8) { e[0] = x0; e[1] = x1; }
  }
  return e[which];
}

static bool build_wps_partition(const std::vector<int>& T, int hop, int win, int grid, std::vector<int>& tsum, std::vector<int>& cut) {
  const int B = (int)T.size();
  const int warm = (win - 1) / hop, minrun = warm + 1;
  tsum.assign(B + 1, 0);
  long long total = 0;
  for (int u = 0; u < B; ++u) {
    total += T[u];
    if (total > 0x3fffffff) return false;
    tsum[u + 1] = (int)total;
  }
  const int G = (int)total;
  const int nw = grid * kWpsWarps;
  // fewer frames per warp than this and the tile kernel serves the batch (TTSA_WPS_MINFRAMES overrides it for experiments;
  // whatever the threshold, the checks at the end reject a partition with a run too short for the hand-over)
  int min_frames = minrun + 1;     // measured: 32 / 40 / 48 x 482 run 0.072 / 0.087 / 0.097 ms here against 0.086 / 0.102 / 0.118 ms in the tile kernel
  if (const char* mf = std::getenv("TTSA_WPS_MINFRAMES")) min_frames = std::max(minrun, std::atoi(mf));
  if (grid <= 0 || (long long)G < (long long)nw * min_frames) return false;
  cut.assign(nw + 1, 0);
  for (int i = 0; i <= nw; ++i) cut[i] = (int)((long long)i * G / nw);
  // snap to utterance boundaries, keep the list monotone
  std::vector<char> on_boundary(nw + 1, 0);
  on_boundary[0] = on_boundary[nw] = 1;
  int u = 0;
  for (int i = 1; i < nw; ++i) {
    int x = cut[i];
    while (u + 1 < B && tsum[u + 1] <= x) ++u;
    if (x - tsum[u] < minrun) x = tsum[u];
    else if (tsum[u + 1] - x < minrun) x = tsum[u + 1];
    cut[i] = x < cut[i - 1] ? cut[i - 1] : x;
  }
  u = 0;
  for (int i = 1; i < nw; ++i) {
    while (u + 1 < B && tsum[u + 1] <= cut[i]) ++u;
    on_boundary[i] = cut[i] == tsum[u];
  }
  // re-space the cuts strictly between two boundary cuts that delimit (part of) ONE utterance
  for (int ia = 0; ia < nw;) {
    int ib = ia + 1;
    while (!on_boundary[ib]) ++ib;
    const int a = cut[ia], b = cut[ib], n = ib - ia;
    if (n > 1 && b > a) {
      int ua = 0;
      { int lo = 0, hi = B; while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (tsum[mid] <= a) lo = mid; else hi = mid; } ua = lo; }
      if (b <= tsum[ua + 1] && (long long)(b - a) >= (long long)n * minrun) {
        // The runs that hold an utterance's first frames (reflected spans, window sums over the frames that exist) and its
        // last ones (the same, plus the flush of the whole window) cost somewhat more than an interior run of the same
        // length; e0 / e1 (in frame times, TTSA_WPS_EDGE) shorten them.  With 30 848 frames over 2 368 warps some run has 14
        // frames whatever the weights, and that run paces the launch: the weights made no measurable difference (default 0).
        const double e0 = (a == tsum[ua]) ? wps_edge_cost(0) : 0.0, e1 = (b == tsum[ua + 1]) ? wps_edge_cost(1) : 0.0;
        const double c = ((double)(b - a) + e0 + e1) / n;            // cost per run, in frames
        if (c - e0 >= minrun && c - e1 >= minrun && c >= minrun + 1) {
          for (int i = ia + 1; i < ib; ++i) cut[i] = a + (int)std::floor((double)(i - ia) * c - e0 + 0.5);
        } else {
          for (int i = ia + 1; i < ib; ++i) cut[i] = a + (int)((long long)(i - ia) * (b - a) / n);
        }
      }
    }
    ia = ib;
  }
  // verify: two cuts strictly inside the same utterance are at least minrun apart, and minrun away from its ends
  u = 0;
  for (int i = 1; i <= nw; ++i) {
    if (cut[i] < cut[i - 1]) return false;
    if (cut[i] == cut[i - 1]) continue;
    while (u + 1 < B && tsum[u + 1] <= cut[i - 1]) ++u;
    const bool prev_inside = cut[i - 1] > tsum[u];
    const bool cur_inside_same = cut[i] < tsum[u + 1];
    if (prev_inside && cur_inside_same && cut[i] - cut[i - 1] < minrun) return false;
    if (prev_inside && cut[i - 1] - tsum[u] < minrun) return false;
    if (cur_inside_same && tsum[u + 1] - cut[i] < minrun) return false;
  }
  return true;
}

static int batch_finish(const ttsa_plan* plan, ttsa_batch* b, ttsa_batch** out, long long frame_stride = 0) {
  const int B = b->B;
  b->device = plan->device;
  b->hop = plan->cfg.hop_length;
  b->frame_off.assign(B + 1, 0); b->wav_off.assign(B + 1, 0); b->tile_off.assign(B + 1, 0); b->chunk_off.assign(B + 1, 0);
  b->max_chunks = 0;
  for (int u = 0; u < B; ++u) {
    if (frame_stride > 0 && b->T[u] > frame_stride) { delete b; return fail(TTSA_ERR_BAD_ARG, "frame_stride %lld < frame count %d of utterance %d", frame_stride, b->T[u], u); }
    b->frame_off[u + 1] = b->frame_off[u] + (frame_stride > 0 ? frame_stride : (long long)b->T[u]);
    b->wav_off[u + 1] = b->wav_off[u] + (b->wav_len[u] + 3) / 4 * 4;
    const long long tiles = (long long)b->tile_off[u] + (b->T[u] + kNF - 1) / kNF;
    if (tiles > std::numeric_limits<int>::max()) { delete b; return fail(TTSA_ERR_BAD_ARG, "batch too large"); }
    b->tile_off[u + 1] = (int)tiles;
    const int ch = (b->wav_len[u] + kDeChunk - 1) / kDeChunk;
    b->chunk_off[u + 1] = b->chunk_off[u] + ch;
    b->max_chunks = std::max(b->max_chunks, ch);
  }
  b->total_frames = b->frame_off[B];
  b->dense = frame_stride <= 0;
  b->total_samples = b->wav_off[B];
  std::memset(&b->dev, 0, sizeof(b->dev));
  b->dev.B = B;
  b->dev.total_tiles = b->tile_off[B];
  {  // fine segments (frame_kernel<..., FINE>): the first 8 frames of an utterance, then 8 - nwarm owned frames each
    const int nwarm = (plan->cfg.win_length - 1) / plan->cfg.hop_length, own = kNF - nwarm;
    b->fine_off.assign(B + 1, 0);
    long long acc = 0;
    for (int u = 0; u < B; ++u) {
      const int T = b->T[u];
      acc += T <= 0 ? 0 : (T <= kNF || own <= 0 ? 1 : 1 + (T - kNF + own - 1) / own);
      b->fine_off[u + 1] = (int)std::min<long long>(acc, std::numeric_limits<int>::max());
    }
    b->dev.total_fine = own >= 2 ? b->fine_off[B] : 0;             // 0: the fine partition does not apply to this geometry
  }
  if (plan->device >= 0 && plan->gl_stream) {
    b->wps_grid = plan->wps_grid;
    b->wps_win = plan->cfg.win_length;
    b->wps_ok = build_wps_partition(b->T, b->hop, b->wps_win, b->wps_grid, b->tsum, b->wps_cut);
    if (b->wps_ok) {                                       // utterance that holds the first frame of each warp's range
      const int nw = (int)b->wps_cut.size() - 1;
      b->wps_u0.assign(nw, 0);
      int u = 0;
      for (int i = 0; i < nw; ++i) {
        while (u + 1 < B && b->tsum[u + 1] <= b->wps_cut[i]) ++u;
        b->wps_u0[i] = u;
      }
    }
  }
  if (plan->device >= 0) {
    DeviceGuard guard(plan->device);
    {
      long long acc = 0;
      b->tsum_all.assign(B + 1, 0);
      for (int u = 0; u < B && acc >= 0; ++u) {
        acc += b->T[u];
        if (acc > 0x3fffffff) acc = -1; else b->tsum_all[u + 1] = (int)acc;
      }
      if (acc < 0) b->tsum_all.clear();
      b->feat_frames = acc > 0 ? (int)acc : 0;
    }
    const size_t n_wps = b->wps_ok ? b->tsum.size() + b->wps_cut.size() + b->wps_u0.size() : 0;
    const size_t n_i = (size_t)B * 2 + (size_t)(B + 1) * 3 + n_wps + b->tsum_all.size();   // T, wav_len, tile_off, chunk_off, [tsum, wps_cut, wps_u0], fine_off, [tsum_all]
    const size_t bytes_i = (n_i * 4 + 15) / 16 * 16;
    const size_t bytes_l = (size_t)(B + 1) * 2 * 8;
    if (cudaMalloc(&b->d_block, bytes_i + bytes_l) != cudaSuccess) { delete b; return fail(TTSA_ERR_CUDA, "cudaMalloc for batch layout failed"); }
    std::vector<char> h(bytes_i + bytes_l, 0);
    long long* hl = (long long*)h.data();
    std::memcpy(hl, b->frame_off.data(), (B + 1) * 8);
    std::memcpy(hl + (B + 1), b->wav_off.data(), (B + 1) * 8);
    int* hi = (int*)(h.data() + bytes_l);
    std::memcpy(hi, b->T.data(), B * 4);
    std::memcpy(hi + B, b->wav_len.data(), B * 4);
    std::memcpy(hi + 2 * B, b->tile_off.data(), (B + 1) * 4);
    std::memcpy(hi + 2 * B + (B + 1), b->chunk_off.data(), (B + 1) * 4);
    if (b->wps_ok) {
      std::memcpy(hi + 2 * B + 2 * (B + 1), b->tsum.data(), b->tsum.size() * 4);
      std::memcpy(hi + 2 * B + 3 * (B + 1), b->wps_cut.data(), b->wps_cut.size() * 4);
      std::memcpy(hi + 2 * B + 3 * (B + 1) + b->wps_cut.size(), b->wps_u0.data(), b->wps_u0.size() * 4);
    }
    const size_t fine_at = 2 * (size_t)B + 2 * (size_t)(B + 1) + n_wps;
    std::memcpy(hi + fine_at, b->fine_off.data(), (B + 1) * 4);
    if (!b->tsum_all.empty()) std::memcpy(hi + fine_at + (B + 1), b->tsum_all.data(), (B + 1) * 4);
    cudaError_t e = cudaMemcpy(b->d_block, h.data(), h.size(), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { cudaFree(b->d_block); delete b; return fail(TTSA_ERR_CUDA, "batch upload: %s", cudaGetErrorString(e)); }
    const long long* dl = (const long long*)b->d_block;
    const int* di = (const int*)((char*)b->d_block + bytes_l);
    b->dev.frame_off = dl;
    b->dev.wav_off = dl + (B + 1);
    b->dev.T = di;
    b->dev.wav_len = di + B;
    b->dev.tile_off = di + 2 * B;
    b->dev.fine_off = di + fine_at;
    b->dev.tsum = b->tsum_all.empty() ? nullptr : di + fine_at + (B + 1);
    b->d_chunk_off = di + 2 * B + (B + 1);
    if (b->wps_ok) {
      b->wps_dev.tsum = di + 2 * B + 2 * (B + 1);
      b->wps_dev.cut = di + 2 * B + 3 * (B + 1);
      b->wps_dev.u0 = b->wps_dev.cut + b->wps_cut.size();
    }
  }
  *out = b;
  return TTSA_OK;
}

extern "C" int ttsa_batch_from_frames(const ttsa_plan* plan, const int32_t* n_frames_host, int32_t n_utts, ttsa_batch** out) {
  return ttsa_batch_from_frames_strided(plan, n_frames_host, n_utts, 0, out);
}

extern "C" int ttsa_batch_from_frames_strided(const ttsa_plan* plan, const int32_t* n_frames_host, int32_t n_utts,
                                              int64_t frame_stride, ttsa_batch** out) {
  if (!plan || !n_frames_host || !out || n_utts <= 0 || frame_stride < 0) return fail(TTSA_ERR_BAD_ARG, "bad argument");
  *out = nullptr;
  ttsa_batch* b = new ttsa_batch();
  b->B = n_utts;
  b->T.resize(n_utts); b->wav_len.resize(n_utts);
  for (int u = 0; u < n_utts; ++u) {
    if (n_frames_host[u] < 0) { delete b; return fail(TTSA_ERR_BAD_ARG, "negative frame count at %d", u); }
    b->T[u] = n_frames_host[u];
    b->wav_len[u] = n_frames_host[u] > 0 ? plan->cfg.hop_length * (n_frames_host[u] - 1) : 0;
  }
  return batch_finish(plan, b, out, frame_stride);
}

extern "C" int ttsa_batch_from_wav_lengths(const ttsa_plan* plan, const int32_t* wav_len_host, int32_t n_utts, ttsa_batch** out) {
  return ttsa_batch_from_wav_lengths_strided(plan, wav_len_host, n_utts, 0, out);
}

extern "C" int ttsa_batch_from_wav_lengths_strided(const ttsa_plan* plan, const int32_t* wav_len_host, int32_t n_utts,
                                                   int64_t frame_stride, ttsa_batch** out) {
  if (!plan || !wav_len_host || !out || n_utts <= 0 || frame_stride < 0) return fail(TTSA_ERR_BAD_ARG, "bad argument");
  *out = nullptr;
  ttsa_batch* b = new ttsa_batch();
  b->B = n_utts;
  b->T.resize(n_utts); b->wav_len.resize(n_utts);
  for (int u = 0; u < n_utts; ++u) {
    if (wav_len_host[u] < 1) { delete b; return fail(TTSA_ERR_BAD_ARG, "wav length at %d must be >= 1 (reflect padding)", u); }
    b->wav_len[u] = wav_len_host[u];
    b->T[u] = 1 + wav_len_host[u] / plan->cfg.hop_length;
  }
  return batch_finish(plan, b, out, frame_stride);
}

extern "C" int ttsa_batch_destroy(ttsa_batch* batch) {
  if (!batch) return TTSA_OK;
  if (batch->d_block) { DeviceGuard g(batch->device); cudaFree(batch->d_block); }
  delete batch;
  return TTSA_OK;
}

extern "C" int64_t ttsa_batch_total_frames(const ttsa_batch* b) { return b ? b->total_frames : -1; }
extern "C" int64_t ttsa_batch_total_samples(const ttsa_batch* b) { return b ? b->total_samples : -1; }

extern "C" int ttsa_batch_offsets(const ttsa_batch* b, int64_t* frame_off, int64_t* wav_off, int32_t* wav_len) {
  if (!b) return fail(TTSA_ERR_BAD_ARG, "null batch");
  for (int u = 0; u <= b->B; ++u) {
    if (frame_off) frame_off[u] = b->frame_off[u];
    if (wav_off) wav_off[u] = b->wav_off[u];
    if (wav_len && u < b->B) wav_len[u] = b->wav_len[u];
  }
  return TTSA_OK;
}

// ---------------------------------------------------------------------------------------------------------
// work calls
// ---------------------------------------------------------------------------------------------------------
static int check_work(const ttsa_plan* plan, const ttsa_batch* batch) {
  if (!plan || !batch) return fail(TTSA_ERR_BAD_ARG, "null plan or batch");
  if (plan->device < 0) return fail(TTSA_ERR_NO_DEVICE, "host-only plan: no CUDA device bound; this library has no CPU path");
  if (batch->device != plan->device) return fail(TTSA_ERR_BAD_ARG, "batch belongs to device %d, plan to %d", batch->device, plan->device);
  if (batch->hop != plan->cfg.hop_length) return fail(TTSA_ERR_BAD_ARG, "batch was laid out for hop %d", batch->hop);
  return TTSA_OK;
}

static int launch_frames(const ttsa_plan* plan, const ttsa_batch* batch, int mode, int src, bool sc,
                         const FrameArgs& args, cudaStream_t st, bool mom = false) {
  if (batch->dev.total_tiles == 0) return TTSA_OK;
  const int max_ctas = plan->ctas_per_sm * plan->num_sms;
  int grid = batch->dev.total_tiles < max_ctas ? batch->dev.total_tiles : max_ctas;
  // small batches (the server's one sentence at a time): one CTA per fine segment, one tile phase per iteration
  const bool fine = mode == MODE_GL_ITER && !mom && !sc && plan->fine_ok && plan->nz == 20 && batch->dev.total_fine > 0 &&
                    batch->dev.total_fine <= max_ctas;
  if (fine) grid = batch->dev.total_fine;
  const size_t smem = ((size_t)plan->geo.ly.sm_total + ((mode == MODE_ANALYSIS && src == OUT_FEATURES) ? plan->geo.mel_smem_floats : 0)) * 4;
  const char* err = launch_frame_kernel(mode, src, plan->nz, sc, plan->fixed_geo, mom, grid, smem, st,
                                        plan->geo, plan->tb, batch->dev, args, fine);
  if (err) return fail(TTSA_ERR_CUDA, "frame kernel launch (mode %d): %s", mode, err);
  return TTSA_OK;
}

extern "C" int ttsa_stft_features(const ttsa_plan* plan, const ttsa_batch* batch, const float* wav_dev,
                                  float* lin_out_dev, float* mel_out_dev, uint32_t flags, void* stream) {
  if (int rc = check_work(plan, batch)) return rc;
  if (!wav_dev || (!lin_out_dev && !mel_out_dev)) return fail(TTSA_ERR_BAD_ARG, "null buffer");
  if ((flags & TTSA_FEAT_PREEMPHASIS) && plan->cfg.preemphasis == 0.0)
    return fail(TTSA_ERR_BAD_CONFIG, " !! Preemphasis is applied with factor 0.0. ");   // utils/audio.py:129-130
  DeviceGuard guard(plan->device);
  FrameArgs a{};
  a.wav_in = wav_dev; a.lin_out = lin_out_dev; a.mel_out = mel_out_dev;
  a.preemph = (flags & TTSA_FEAT_PREEMPHASIS) ? 1 : 0;
  if (plan->generic) return gen_launch(plan, batch, MODE_ANALYSIS, OUT_FEATURES, a, (cudaStream_t)stream);
  // warp-stream kernel from two frames per warp of the GPU upwards (TTSA_FEAT_MINFRAMES overrides the threshold,
  // TTSA_FEAT_KERNEL=tile the choice); the tile kernel serves smaller batches and run-time geometries
  if (plan->feat_stream && plan->fixed_geo && batch->dev.tsum != nullptr) {
    int min_frames = 2;
    if (const char* mf = std::getenv("TTSA_FEAT_MINFRAMES")) min_frames = std::max(0, std::atoi(mf));
    if ((long long)batch->feat_frames >= (long long)min_frames * plan->wps_grid * kWpsWarps && batch->feat_frames > 0) {
      const char* err = launch_feat_stream(plan->cfg.hop_length, plan->cfg.win_length, plan->wps_grid, (cudaStream_t)stream, plan->geo,
                                           plan->tb, batch->dev, a, batch->feat_frames);
      if (err) return fail(TTSA_ERR_CUDA, "feature stream kernel launch: %s", err);
      return TTSA_OK;
    }
  }
  return launch_frames(plan, batch, MODE_ANALYSIS, OUT_FEATURES, false, a, (cudaStream_t)stream);
}

extern "C" int ttsa_stft(const ttsa_plan* plan, const ttsa_batch* batch, const float* wav_dev, float* stft_out_dev, void* stream) {
  if (int rc = check_work(plan, batch)) return rc;
  if (!wav_dev || !stft_out_dev) return fail(TTSA_ERR_BAD_ARG, "null buffer");
  DeviceGuard guard(plan->device);
  FrameArgs a{};
  a.wav_in = wav_dev; a.cplx_out = stft_out_dev;
  if (plan->generic) return gen_launch(plan, batch, MODE_ANALYSIS, OUT_COMPLEX, a, (cudaStream_t)stream);
  return launch_frames(plan, batch, MODE_ANALYSIS, OUT_COMPLEX, false, a, (cudaStream_t)stream);
}

extern "C" int ttsa_istft(const ttsa_plan* plan, const ttsa_batch* batch, const float* stft_dev, float* wav_out_dev, void* stream) {
  if (int rc = check_work(plan, batch)) return rc;
  if (!stft_dev || !wav_out_dev) return fail(TTSA_ERR_BAD_ARG, "null buffer");
  DeviceGuard guard(plan->device);
  FrameArgs a{};
  a.cplx_in = stft_dev; a.wav_out = wav_out_dev; a.rows_total = batch->total_frames;
  if (plan->generic) return gen_synthesise(plan, batch, MODE_SYNTH, SRC_COMPLEX, a, (cudaStream_t)stream);
  return launch_frames(plan, batch, MODE_SYNTH, SRC_COMPLEX, false, a, (cudaStream_t)stream);
}

extern "C" size_t ttsa_deemphasis_workspace_bytes(const ttsa_plan* plan, const ttsa_batch* batch) {
  if (!plan || !batch) return 0;
  return ((size_t)batch->chunk_off[batch->B] * 4 + 255) / 256 * 256 + 256;
}

static size_t wps_flag_bytes(const ttsa_batch* batch) {
  size_t n = batch->wps_ok ? 2 * (((size_t)batch->wps_grid * kWpsWarps * 4 + 255) / 256 * 256) : 0;   // head-zone flags, done flags
#ifdef TTSA_WPS_TRACE
  if (batch->wps_ok) n += (size_t)batch->wps_grid * kWpsWarps * 64 * 8 * 8;                          // stamps, at the very end
#endif
  return n;
}

extern "C" size_t ttsa_griffin_lim_workspace_bytes(const ttsa_plan* plan, const ttsa_batch* batch) {
  if (!plan || !batch) return 0;
  const size_t wav = ((size_t)batch->total_samples * 4 + 255) / 256 * 256 + 256;
  // two estimates + one more for the momentum mode, the de-emphasis aggregates, the warp flags of the stream kernel
  return 3 * wav + ttsa_deemphasis_workspace_bytes(plan, batch) + wps_flag_bytes(batch);
}

static int deemph_launch(const ttsa_plan* plan, const ttsa_batch* batch, const float* x, float* y, float* agg, cudaStream_t st) {
  if (batch->max_chunks == 0) return TTSA_OK;
  DeParams dp;
  dp.p = (float)plan->cfg.preemphasis;
  dp.log2p = (float)std::log2(plan->cfg.preemphasis);
  dp.chunk_off = batch->d_chunk_off;
  dim3 grid(batch->max_chunks, batch->B);
  deemph_aggregate_kernel<<<grid, kDeThreads, 0, st>>>(batch->dev, dp, x, agg);
  deemph_apply_kernel<<<grid, kDeThreads, 0, st>>>(batch->dev, dp, agg, x, y);
  g_launches += 2;
  CUDA_TRY(cudaGetLastError());
  return TTSA_OK;
}

extern "C" int ttsa_deemphasis(const ttsa_plan* plan, const ttsa_batch* batch, const float* x_dev, float* y_dev,
                               void* workspace_dev, size_t workspace_bytes, void* stream) {
  if (int rc = check_work(plan, batch)) return rc;
  if (!x_dev || !y_dev || !workspace_dev) return fail(TTSA_ERR_BAD_ARG, "null buffer");
  if (plan->cfg.preemphasis == 0.0) return fail(TTSA_ERR_BAD_CONFIG, " !! Preemphasis is applied with factor 0.0. ");
  if (plan->cfg.preemphasis < 0.0) return fail(TTSA_ERR_UNSUPPORTED, "negative preemphasis");
  if (workspace_bytes < ttsa_deemphasis_workspace_bytes(plan, batch)) return fail(TTSA_ERR_WORKSPACE, "workspace too small");
  DeviceGuard guard(plan->device);
  return deemph_launch(plan, batch, x_dev, y_dev, (float*)workspace_dev, (cudaStream_t)stream);
}

// ---------------------------------------------------------------------------------------------------------
// waveform post-processing
// ---------------------------------------------------------------------------------------------------------
static dim3 wav_grid(const ttsa_batch* batch, long long extra = 0) {
  long long maxlen = 0;
  for (int v : batch->wav_len) maxlen = std::max<long long>(maxlen, v);
  maxlen += extra;
  return dim3((unsigned)std::max<long long>(1, std::min<long long>(1024, (maxlen + 1023) / 1024)), (unsigned)batch->B);
}

extern "C" int ttsa_wav_peaks(const ttsa_plan* plan, const ttsa_batch* batch, const float* wav_dev, const int32_t* lens_dev,
                              float* peaks_dev, void* stream) {
  if (int rc = check_work(plan, batch)) return rc;
  if (!wav_dev || !peaks_dev) return fail(TTSA_ERR_BAD_ARG, "null buffer");
  if (batch->B == 0) return TTSA_OK;
  DeviceGuard guard(plan->device);
  cudaStream_t st = (cudaStream_t)stream;
  CUDA_TRY(cudaMemsetAsync(peaks_dev, 0, (size_t)batch->B * 4, st));
  wav_peak_kernel<<<wav_grid(batch), 256, 0, st>>>(batch->dev, lens_dev, wav_dev, (unsigned*)peaks_dev);
  g_launches += 1;
  CUDA_TRY(cudaGetLastError());
  return TTSA_OK;
}

extern "C" int ttsa_find_endpoint(const ttsa_plan* plan, const ttsa_batch* batch, const float* wav_dev, double threshold_db,
                                  double min_silence_sec, int32_t* endpoints_dev, void* stream) {
  if (int rc = check_work(plan, batch)) return rc;
  if (!wav_dev || !endpoints_dev) return fail(TTSA_ERR_BAD_ARG, "null buffer");
  const int window = (int)(plan->cfg.sample_rate * min_silence_sec);
  const int hop = window / 4;
  if (window < 4) return fail(TTSA_ERR_BAD_ARG, "min_silence_sec %g gives a window of %d samples", min_silence_sec, window);
  if (batch->B == 0) return TTSA_OK;
  DeviceGuard guard(plan->device);
  cudaStream_t st = (cudaStream_t)stream;
  wav_endpoint_init_kernel<<<(batch->B + 255) / 256, 256, 0, st>>>(batch->dev, endpoints_dev);
  g_launches += 1;
  long long maxlen = 0;
  for (int v : batch->wav_len) maxlen = std::max<long long>(maxlen, v);
  // candidates x = hop, 2 hop, ... < L - window
  const long long ncand = maxlen - window > hop ? (maxlen - window - hop + hop - 1) / hop : 0;
  if (ncand > 0) {
    const double threshold = std::pow(10.0, threshold_db * 0.05);          // _db_to_amp, utils/audio.py:125-126
    wav_endpoint_kernel<<<dim3((unsigned)ncand, (unsigned)batch->B), 256, 0, st>>>(batch->dev, wav_dev, window, hop, threshold,
                                                                                  endpoints_dev);
    g_launches += 1;
  }
  CUDA_TRY(cudaGetLastError());
  return TTSA_OK;
}

extern "C" size_t ttsa_pcm16_workspace_bytes(const ttsa_plan* plan, const ttsa_batch* batch) {
  if (!plan || !batch) return 0;
  return ((size_t)batch->B * 8 + 255) / 256 * 256 + ((size_t)batch->B * 4 + 255) / 256 * 256 + 256;
}

extern "C" int ttsa_wav_to_pcm16(const ttsa_plan* plan, const ttsa_batch* batch, const float* wav_dev, const int32_t* lens_dev,
                                 uint32_t flags, int64_t gap_samples, int64_t* out_off_dev, int16_t* out_dev,
                                 int64_t out_capacity, void* workspace_dev, size_t workspace_bytes, void* stream) {
  if (int rc = check_work(plan, batch)) return rc;
  if (!wav_dev || !out_off_dev || !out_dev || !workspace_dev) return fail(TTSA_ERR_BAD_ARG, "null buffer");
  if (gap_samples < 0 || out_capacity < 0) return fail(TTSA_ERR_BAD_ARG, "negative gap or capacity");
  if (workspace_bytes < ttsa_pcm16_workspace_bytes(plan, batch)) return fail(TTSA_ERR_WORKSPACE, "workspace too small");
  long long sum_len = 0;
  for (int v : batch->wav_len) sum_len += v;
  if (lens_dev == nullptr && out_capacity < sum_len + gap_samples * batch->B)
    return fail(TTSA_ERR_BAD_ARG, "out_capacity %lld < %lld samples", (long long)out_capacity,
                (long long)(sum_len + gap_samples * batch->B));
  if (batch->B == 0) return TTSA_OK;
  DeviceGuard guard(plan->device);
  cudaStream_t st = (cudaStream_t)stream;
  double* scales = (double*)workspace_dev;
  unsigned* peaks = (unsigned*)((char*)workspace_dev + ((size_t)batch->B * 8 + 255) / 256 * 256);
  CUDA_TRY(cudaMemsetAsync(peaks, 0, (size_t)batch->B * 4, st));
  wav_peak_kernel<<<wav_grid(batch), 256, 0, st>>>(batch->dev, lens_dev, wav_dev, peaks);
  pcm_plan_kernel<<<1, 256, 0, st>>>(batch->dev, lens_dev, peaks, (flags & TTSA_PCM_JOINT_PEAK) ? 1 : 0, (long long)gap_samples,
                                     (long long*)out_off_dev, scales);
  pcm_convert_kernel<<<wav_grid(batch, gap_samples), 256, 0, st>>>(batch->dev, lens_dev, wav_dev, (const long long*)out_off_dev,
                                                                  scales, (long long)gap_samples,
                                                                  (flags & TTSA_PCM_F32_ARITH) ? 1 : 0, (long long)out_capacity,
                                                                  (short*)out_dev);
  g_launches += 3;
  CUDA_TRY(cudaGetLastError());
  return TTSA_OK;
}

extern "C" int ttsa_preemphasis(const ttsa_plan* plan, const ttsa_batch* batch, const float* x_dev, float* y_dev, void* stream) {
  if (int rc = check_work(plan, batch)) return rc;
  if (!x_dev || !y_dev || x_dev == y_dev) return fail(TTSA_ERR_BAD_ARG, "null or aliasing buffers");
  if (plan->cfg.preemphasis == 0.0) return fail(TTSA_ERR_BAD_CONFIG, " !! Preemphasis is applied with factor 0.0. ");
  DeviceGuard guard(plan->device);
  int maxlen = 0;
  for (int v : batch->wav_len) maxlen = std::max(maxlen, v);
  dim3 grid(std::max(1, std::min(1024, (maxlen + 255) / 256)), batch->B);
  preemphasis_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(batch->dev, (float)plan->cfg.preemphasis, x_dev, y_dev);
  g_launches += 1;
  CUDA_TRY(cudaGetLastError());
  return TTSA_OK;
}

extern "C" int ttsa_griffin_lim(const ttsa_plan* plan, const ttsa_batch* batch, const float* spec_dev, int spec_kind,
                                int iters, const float* init_angles_dev, uint64_t seed, uint32_t flags,
                                float* wav_out_dev, float* sc_log_dev, void* workspace_dev, size_t workspace_bytes,
                                void* stream) {
  return ttsa_griffin_lim_fast(plan, batch, spec_dev, spec_kind, iters, init_angles_dev, seed, flags, 0.0, wav_out_dev,
                               sc_log_dev, workspace_dev, workspace_bytes, stream);
}

extern "C" int ttsa_griffin_lim_fast(const ttsa_plan* plan, const ttsa_batch* batch, const float* spec_dev, int spec_kind,
                                     int iters, const float* init_angles_dev, uint64_t seed, uint32_t flags, double momentum,
                                     float* wav_out_dev, float* sc_log_dev, void* workspace_dev, size_t workspace_bytes,
                                     void* stream) {
  if (int rc = check_work(plan, batch)) return rc;
  if (!(momentum >= 0.0 && momentum < 1.0)) return fail(TTSA_ERR_BAD_ARG, "momentum %g outside [0, 1)", momentum);
  if (!spec_dev || !wav_out_dev || !workspace_dev) return fail(TTSA_ERR_BAD_ARG, "null buffer");
  if (spec_kind != TTSA_SPEC_MAGNITUDE && spec_kind != TTSA_SPEC_NORM_DB) return fail(TTSA_ERR_BAD_ARG, "bad spec_kind %d", spec_kind);
  if (iters < 0) return fail(TTSA_ERR_BAD_ARG, "iters < 0");
  // waveform spans are read as sample pairs from 16-byte-aligned utterance slots (spectrogram rows may start anywhere)
  if (((uintptr_t)wav_out_dev | (uintptr_t)workspace_dev) & 15u)
    return fail(TTSA_ERR_BAD_ARG, "wav_out and workspace must be 16-byte aligned");
  if (workspace_bytes < ttsa_griffin_lim_workspace_bytes(plan, batch)) return fail(TTSA_ERR_WORKSPACE, "workspace too small: %zu < %zu", workspace_bytes, ttsa_griffin_lim_workspace_bytes(plan, batch));
  const bool deemph = (flags & TTSA_GL_DEEMPHASIS) != 0;
  if (deemph && plan->cfg.preemphasis == 0.0) return fail(TTSA_ERR_BAD_CONFIG, " !! Preemphasis is applied with factor 0.0. ");
  if (deemph && plan->cfg.preemphasis < 0.0) return fail(TTSA_ERR_UNSUPPORTED, "de-emphasis with a negative coefficient (%g) is not supported", plan->cfg.preemphasis);
  DeviceGuard guard(plan->device);
  cudaStream_t st = (cudaStream_t)stream;
  const size_t wav_bytes = ((size_t)batch->total_samples * 4 + 255) / 256 * 256 + 256;
  float* wsA = (float*)workspace_dev;
  float* wsB = (float*)((char*)workspace_dev + wav_bytes);
  float* wsC = (float*)((char*)workspace_dev + 2 * wav_bytes);
  float* agg = (float*)((char*)workspace_dev + 3 * wav_bytes);
  int* wps_flags = (int*)((char*)agg + ttsa_deemphasis_workspace_bytes(plan, batch));
  // write k (0 = initial istft, 1..iters = iterations) goes to bufs[k % nb]; the last one must land in `last`.
  // nb = 2 (ping-pong), or 3 with momentum: iteration k reads estimates k-1 and k-2.
  const bool mom = momentum > 0.0 && iters > 1;
  const float beta = (float)(momentum / (1.0 + momentum));
  const int nb = mom ? 3 : 2;
  float* bufs[3] = {nullptr, nullptr, nullptr};
  float* last;
  if (deemph) { bufs[0] = wsA; bufs[1] = wsB; bufs[2] = wsC; last = bufs[iters % nb]; }
  else {
    last = wav_out_dev;
    bufs[iters % nb] = wav_out_dev;
    float* spare[2] = {wsA, wsB};
    for (int i = 0, k = 0; i < nb; ++i) if (bufs[i] == nullptr) bufs[i] = spare[k++];
  }
  if (sc_log_dev && iters > 0) CUDA_TRY(cudaMemsetAsync(sc_log_dev, 0, (size_t)iters * batch->B * 2 * 4, st));
  const bool use_stream = !plan->generic && plan->gl_stream && plan->fixed_geo && plan->debug == 0 && batch->wps_ok &&
                          batch->wps_win == plan->cfg.win_length;
  if (use_stream && iters > 0) CUDA_TRY(cudaMemsetAsync(wps_flags, 0, wps_flag_bytes(batch), st));

  if (plan->generic) {
    FrameArgs a{};
    a.spec = spec_dev; a.angles = init_angles_dev; a.seed = seed; a.wav_out = bufs[0];
    if (int rc = gen_synthesise(plan, batch, MODE_SYNTH, spec_kind, a, st)) return rc;
    for (int i = 1; i <= iters; ++i) {
      FrameArgs b{};
      b.spec = spec_dev; b.wav_in = bufs[(i - 1) % nb]; b.wav_out = bufs[i % nb];
      if (mom && i >= 2) { b.wav_prev = bufs[(i - 2) % nb]; b.beta = beta; }
      b.sc_acc = sc_log_dev ? sc_log_dev + (size_t)(i - 1) * batch->B * 2 : nullptr;
      if (int rc = gen_synthesise(plan, batch, MODE_GL_ITER, spec_kind, b, st)) return rc;
    }
    if (bufs[iters % nb] != last) return fail(TTSA_ERR_CUDA, "internal: buffer rotation");
    if (deemph) return deemph_launch(plan, batch, last, wav_out_dev, agg, st);
    return TTSA_OK;
  }
  FrameArgs a{};
  a.spec = spec_dev; a.spec_end = spec_dev + (size_t)batch->total_frames * kF; a.rows_total = batch->total_frames;
  a.angles = init_angles_dev; a.seed = seed; a.wav_out = bufs[0];
  if (int rc = launch_frames(plan, batch, MODE_SYNTH, spec_kind, false, a, st)) return rc;
  // The warp-stream kernel can run several iterations in ONE launch (neighbour-only synchronisation between iterations
  // through global flags, gl_stream.cuh): TTSA_GL_FUSE=n runs up to n iterations per launch.  Measured at 64 x 482
  // (DESIGN.md, K3 "fused iterations"): 0.1416 ms per iteration fused against 0.1285 ms with one launch per iteration --
  // the hand-over at a run's end (wait, finish the next run's head zone, publish: 12 us) is on every warp's cycle when
  // iterations are chained, while a kernel boundary overlaps it with the other warps' last frames -- so the default is 1.
  int fuse = 1;
  if (const char* fz = std::getenv("TTSA_GL_FUSE")) fuse = std::max(1, std::atoi(fz));
  for (int i = 1; i <= iters; ++i) {
    FrameArgs b{};
    b.spec = spec_dev; b.spec_end = spec_dev + (size_t)batch->total_frames * kF;
    b.wav_in = bufs[(i - 1) % nb]; b.wav_out = bufs[i % nb];
    const bool mom_i = mom && i >= 2;                         // the first iteration has no estimate before its input
    if (mom_i) { b.wav_prev = bufs[(i - 2) % nb]; b.beta = beta; }
    b.wav_end = b.wav_in + batch->total_samples;
    b.debug = plan->debug;
    b.sc_acc = sc_log_dev ? sc_log_dev + (size_t)(i - 1) * batch->B * 2 : nullptr;
    if (!mom_i && use_stream) {
      b.wps_flags = wps_flags; b.wps_epoch = i;
      b.wps_done = wps_flags + ((size_t)batch->wps_grid * kWpsWarps * 4 + 255) / 256 * 64;
#ifdef TTSA_WPS_TRACE
      b.wps_trace = (unsigned long long*)((char*)wps_flags + wps_flag_bytes(batch) - (size_t)batch->wps_grid * kWpsWarps * 64 * 8 * 8);
#endif
      b.wps_iters = mom ? 1 : std::min(fuse, iters - i + 1);  // nb == 2 without momentum: the kernel's ping-pong is bufs[]'s
      i += b.wps_iters - 1;
      const char* err = launch_gl_stream(spec_kind, sc_log_dev != nullptr, plan->cfg.hop_length, plan->cfg.win_length, batch->wps_grid,
                                         st, plan->geo, plan->tb, batch->dev, batch->wps_dev, b);
      if (err) return fail(TTSA_ERR_CUDA, "Griffin-Lim stream kernel launch: %s", err);
    } else if (int rc = launch_frames(plan, batch, MODE_GL_ITER, spec_kind, sc_log_dev != nullptr, b, st, mom_i)) {
      return rc;
    }
  }
  if (bufs[iters % nb] != last) return fail(TTSA_ERR_CUDA, "internal: buffer rotation");
  if (deemph) return deemph_launch(plan, batch, last, wav_out_dev, agg, st);
  return TTSA_OK;
}

extern "C" int ttsa_mel_to_linear(const ttsa_plan* plan, const ttsa_batch* batch, const float* mel_dev, int in_kind,
                                  float* lin_out_dev, int out_kind, void* stream) {
  if (int rc = check_work(plan, batch)) return rc;
  if (!mel_dev || !lin_out_dev) return fail(TTSA_ERR_BAD_ARG, "null buffer");
  if ((in_kind != 0 && in_kind != 1) || (out_kind != 0 && out_kind != 1)) return fail(TTSA_ERR_BAD_ARG, "bad in/out kind");
  if (batch->total_frames == 0) return TTSA_OK;
  DeviceGuard guard(plan->device);
  MelParams mp = plan->mel;
  mp.rows = batch->total_frames;
  if (plan->generic || plan->mel_gemm == 1) {   // fp32 SIMT kernel (any num_freq; profiling / cross-check)
    dim3 grid((mp.F + kMtlBins - 1) / kMtlBins, (unsigned)((batch->total_frames + kMtlRows - 1) / kMtlRows));
    const size_t smem = (size_t)mp.num_mels * (kMtlBins + kMtlRows) * 4;
    mel_to_linear_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(mp, plan->d_pinvT, plan->ldp, mel_dev, lin_out_dev, in_kind, out_kind);
  } else if (plan->d_pinv_tc128 != nullptr && plan->mel_gemm == 0) {
    // transposed, warp-specialised tensor-core kernel: bins as the UMMA M dimension, coalesced stores from registers
    TcGemmParams tp;
    tp.mp = mp; tp.a = mel_dev; tp.lda = mp.num_mels; tp.k_total = mp.num_mels; tp.n_chunks = 1;
    tp.b = plan->d_pinv_tc128; tp.out = lin_out_dev; tp.ldo = kF; tp.n_valid = kF; tp.in_kind = in_kind; tp.out_kind = out_kind;
    tp.mode = 0;
    const int n_ftiles = (int)((batch->total_frames + kT2Frames - 1) / kT2Frames);
    const int grid = n_ftiles < plan->num_sms ? n_ftiles : plan->num_sms;
    mel_to_linear_tc2_kernel<<<grid, kT2Threads, kT2Smem, (cudaStream_t)stream>>>(tp, n_ftiles);
  } else if (plan->d_pinv_tc96 != nullptr && plan->mel_gemm != 2) {
    // pipelined tensor-core kernel: one CTA per 128-frame tile walks all bin tiles
    TcGemmParams tp;
    tp.mp = mp; tp.a = mel_dev; tp.lda = mp.num_mels; tp.k_total = mp.num_mels; tp.n_chunks = 1;
    tp.b = plan->d_pinv_tc96; tp.out = lin_out_dev; tp.ldo = kF; tp.n_valid = kF; tp.in_kind = in_kind; tp.out_kind = out_kind;
    tp.mode = 0;
    const int n_mtiles = (int)((batch->total_frames + kTcRows - 1) / kTcRows);
    const int grid = n_mtiles < plan->num_sms ? n_mtiles : plan->num_sms;
    mel_to_linear_tc_kernel<<<grid, kM2lThreads, kM2lSmem, (cudaStream_t)stream>>>(tp, n_mtiles);
  } else {
    TcGemmParams tp;
    tp.mp = mp; tp.a = mel_dev; tp.lda = mp.num_mels; tp.k_total = mp.num_mels; tp.n_chunks = plan->pinv_chunks;
    tp.b = plan->d_pinv_tc; tp.out = lin_out_dev; tp.ldo = kF; tp.n_valid = kF; tp.in_kind = in_kind; tp.out_kind = out_kind;
    tp.mode = 0;
    dim3 grid((unsigned)((batch->total_frames + kTcRows - 1) / kTcRows), 5);
    const size_t smem = 3 * kTcRows * kTcChunk * 2 + 3 * 208 * kTcChunk * 2;
    gemm_bf16x3_tc_kernel<208><<<grid, kTcThreads, smem, (cudaStream_t)stream>>>(tp);
  }
  g_launches += 1;
  CUDA_TRY(cudaGetLastError());
  return TTSA_OK;
}

extern "C" int ttsa_linear_to_mel(const ttsa_plan* plan, const ttsa_batch* batch, const float* lin_dev, int in_kind,
                                  float* mel_out_dev, int out_kind, void* stream) {
  if (int rc = check_work(plan, batch)) return rc;
  if (!lin_dev || !mel_out_dev) return fail(TTSA_ERR_BAD_ARG, "null buffer");
  if ((in_kind != 0 && in_kind != 1) || (out_kind != 0 && out_kind != 2)) return fail(TTSA_ERR_BAD_ARG, "bad in/out kind");
  if (batch->total_frames == 0) return TTSA_OK;
  DeviceGuard guard(plan->device);
  MelParams mp = plan->mel;
  mp.rows = batch->total_frames;
  if (plan->generic) {
    const unsigned grid = (unsigned)((batch->total_frames + 7) / 8);
    gen_linear_to_mel_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(mp, mp.F, plan->gt, lin_dev, mel_out_dev, in_kind, out_kind);
  } else if (plan->d_mel_tc == nullptr || plan->mel_gemm == 1) {   // banded fp32 SIMT kernel
    const unsigned grid = (unsigned)((batch->total_frames + 7) / 8);
    linear_to_mel_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(mp, plan->tb, lin_dev, mel_out_dev, in_kind, out_kind);
  } else {
    TcGemmParams tp;
    tp.mp = mp; tp.a = lin_dev; tp.lda = kF; tp.k_total = kF; tp.n_chunks = 13;
    tp.b = plan->d_mel_tc; tp.out = mel_out_dev; tp.ldo = mp.num_mels; tp.n_valid = mp.num_mels; tp.in_kind = in_kind;
    tp.out_kind = out_kind; tp.mode = 1;
    dim3 grid((unsigned)((batch->total_frames + kTcRows - 1) / kTcRows), 1);
    const size_t smem = 3 * kTcRows * kTcChunk * 2 + 3 * 80 * kTcChunk * 2;
    gemm_bf16x3_tc_kernel<80><<<grid, kTcThreads, smem, (cudaStream_t)stream>>>(tp);
  }
  g_launches += 1;
  CUDA_TRY(cudaGetLastError());
  return TTSA_OK;
}

extern "C" int ttsa_pointwise(const ttsa_plan* plan, int op, const float* x_dev, float* y_dev, int64_t n, void* stream) {
  if (!plan) return fail(TTSA_ERR_BAD_ARG, "null plan");
  if (plan->device < 0) return fail(TTSA_ERR_NO_DEVICE, "host-only plan: no CUDA device bound; this library has no CPU path");
  if (op < 0 || op > 3) return fail(TTSA_ERR_BAD_ARG, "bad op %d", op);
  if (n < 0 || (n > 0 && (!x_dev || !y_dev))) return fail(TTSA_ERR_BAD_ARG, "null buffer");
  if (n == 0) return TTSA_OK;
  DeviceGuard guard(plan->device);
  PwParams p = plan->pw;
  p.op = op;
  const long long blocks = (n + 255) / 256;
  const int grid = (int)std::min<long long>(blocks, (long long)plan->num_sms * 16);
  pointwise_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(p, x_dev, y_dev, (long long)n);
  g_launches += 1;
  CUDA_TRY(cudaGetLastError());
  return TTSA_OK;
}

extern "C" int ttsa_transpose(const float* in_dev, float* out_dev, int64_t rows, int64_t cols, void* stream) {
  if (rows < 0 || cols < 0) return fail(TTSA_ERR_BAD_ARG, "negative shape");
  if (rows == 0 || cols == 0) return TTSA_OK;
  if (!in_dev || !out_dev || in_dev == out_dev) return fail(TTSA_ERR_BAD_ARG, "null or aliasing buffers");
  const long long gy = (rows + 31) / 32, gx = (cols + 31) / 32;
  if (gy > 65535) return fail(TTSA_ERR_UNSUPPORTED, "too many rows for one launch");
  dim3 grid((unsigned)gx, (unsigned)gy), block(32, 8);
  transpose_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(in_dev, out_dev, rows, cols);
  g_launches += 1;
  CUDA_TRY(cudaGetLastError());
  return TTSA_OK;
}
