// Host-side (float64) construction of the plan's immutable tables: Hann window, twiddles, window-sum-square,
// Slaney mel basis (librosa.filters.mel semantics, utils/audio.py:68-77) and its Moore-Penrose pseudo-inverse
// (np.linalg.pinv, utils/audio.py:65).  Table construction only -- no signal data passes through here.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

namespace ttsa_host {

constexpr double kPi = 3.14159265358979323846264338327950288;

// scipy.signal.get_window('hann', win, fftbins=True)
inline std::vector<double> hann_periodic(int win) {
  std::vector<double> w(win);
  for (int n = 0; n < win; ++n) w[n] = 0.5 - 0.5 * std::cos(2.0 * kPi * n / win);
  return w;
}

inline double hz_to_mel(double f) {
  const double f_sp = 200.0 / 3.0, min_log_hz = 1000.0, min_log_mel = min_log_hz / f_sp;
  const double logstep = std::log(6.4) / 27.0;
  return f >= min_log_hz ? min_log_mel + std::log(f / min_log_hz) / logstep : f / f_sp;
}

inline double mel_to_hz(double m) {
  const double f_sp = 200.0 / 3.0, min_log_hz = 1000.0, min_log_mel = min_log_hz / f_sp;
  const double logstep = std::log(6.4) / 27.0;
  return m >= min_log_mel ? min_log_hz * std::exp(logstep * (m - min_log_mel)) : f_sp * m;
}

// librosa.filters.mel(sr, n_fft, n_mels, fmin, fmax, htk=False, norm=1): [n_mels][n_bins] row-major
inline std::vector<double> mel_basis(int sr, int n_fft, int n_mels, double fmin, double fmax) {
  const int n_bins = 1 + n_fft / 2;
  std::vector<double> weights((size_t)n_mels * n_bins, 0.0);
  std::vector<double> fftfreqs(n_bins), mel_f(n_mels + 2);
  for (int k = 0; k < n_bins; ++k) fftfreqs[k] = (0.5 * sr) * k / (n_bins - 1);
  const double lo = hz_to_mel(fmin), hi = hz_to_mel(fmax);
  for (int i = 0; i < n_mels + 2; ++i) mel_f[i] = mel_to_hz(lo + (hi - lo) * i / (n_mels + 1));
  for (int i = 0; i < n_mels; ++i) {
    const double fd0 = mel_f[i + 1] - mel_f[i], fd1 = mel_f[i + 2] - mel_f[i + 1];
    const double enorm = 2.0 / (mel_f[i + 2] - mel_f[i]);
    for (int k = 0; k < n_bins; ++k) {
      const double lower = -(mel_f[i] - fftfreqs[k]) / fd0;
      const double upper = (mel_f[i + 2] - fftfreqs[k]) / fd1;
      weights[(size_t)i * n_bins + k] = std::max(0.0, std::min(lower, upper)) * enorm;
    }
  }
  return weights;
}

// pinv(M) for M [m x n], m <= n, through a one-sided (Hestenes) Jacobi SVD of M itself: row pairs are rotated until the
// rows are mutually orthogonal, M = U A with orthogonal rows a_i (||a_i|| = sigma_i), so M^+ = sum_i a_i u_i^T / sigma_i^2
// over sigma_i > rcond * sigma_max with numpy's cut-off (np.linalg.pinv: rcond = 1e-15, utils/audio.py:65).  Working on
// M rather than on the Gram matrix M M^T keeps the condition number unsquared (near rank-deficient bases: many mels
// with a small n_fft, a high fmin).  Returns [n x m].
inline std::vector<double> pinv_wide(const std::vector<double>& M, int m, int n) {
  std::vector<double> A(M), U((size_t)m * m, 0.0);
  for (int i = 0; i < m; ++i) U[(size_t)i * m + i] = 1.0;
  for (int sweep = 0; sweep < 60; ++sweep) {
    bool rotated = false;
    for (int p = 0; p < m - 1; ++p) {
      for (int q = p + 1; q < m; ++q) {
        double alpha = 0.0, beta = 0.0, gamma = 0.0;
        const double* ap = &A[(size_t)p * n];
        const double* aq = &A[(size_t)q * n];
        for (int k = 0; k < n; ++k) { alpha += ap[k] * ap[k]; beta += aq[k] * aq[k]; gamma += ap[k] * aq[k]; }
        if (gamma == 0.0 || std::fabs(gamma) <= 1e-15 * std::sqrt(alpha * beta)) continue;
        rotated = true;
        const double zeta = (beta - alpha) / (2.0 * gamma);
        const double t = (zeta >= 0 ? 1.0 : -1.0) / (std::fabs(zeta) + std::sqrt(1.0 + zeta * zeta));
        const double c = 1.0 / std::sqrt(1.0 + t * t), sn = c * t;
        double* bp = &A[(size_t)p * n];
        double* bq = &A[(size_t)q * n];
        for (int k = 0; k < n; ++k) {
          const double x = bp[k], y = bq[k];
          bp[k] = c * x - sn * y;
          bq[k] = sn * x + c * y;
        }
        for (int k = 0; k < m; ++k) {   // U <- U R^T: the same rotation on columns p, q
          const double x = U[(size_t)k * m + p], y = U[(size_t)k * m + q];
          U[(size_t)k * m + p] = c * x - sn * y;
          U[(size_t)k * m + q] = sn * x + c * y;
        }
      }
    }
    if (!rotated) break;
  }
  std::vector<double> sig2(m, 0.0);
  double smax2 = 0.0;
  for (int i = 0; i < m; ++i) {
    double acc = 0.0;
    for (int k = 0; k < n; ++k) acc += A[(size_t)i * n + k] * A[(size_t)i * n + k];
    sig2[i] = acc;
    smax2 = std::max(smax2, acc);
  }
  const double cut2 = 1e-30 * smax2;            // sigma_i > 1e-15 * sigma_max
  std::vector<double> P((size_t)n * m, 0.0);
  for (int i = 0; i < m; ++i) {
    if (!(sig2[i] > cut2) || sig2[i] == 0.0) continue;
    const double inv = 1.0 / sig2[i];
    for (int k = 0; k < n; ++k) {
      const double a = A[(size_t)i * n + k] * inv;
      if (a == 0.0) continue;
      for (int j = 0; j < m; ++j) P[(size_t)k * m + j] += a * U[(size_t)j * m + i];
    }
  }
  return P;
}

// Segment schedule of a banded (triangular, 50 % overlap) mel basis for the warp-stream feature kernel (feat_stream.cuh).
// A bin between the peaks of filters s and s + 1 has (at most) two taps: a falling one for filter s, a rising one for s + 1.
// The bins are grouped by that s ("segments", s = -1 .. num_mels - 1), long segments are split into pieces, and every piece
// becomes one (slot, lane) CELL, slot = 0..2: the lane reads each of its bins ONCE and feeds two accumulators (D: falling tap
// -> filter s, U: rising tap -> filter s + 1), so a step moves 8 bytes of weights where a per-filter walk moves a tap per
// filter and bin.  The three slots run back to back with trip counts common to the warp (cells sorted by size, so a slot is as
// long as its largest cell); inside a slot the order in which a lane walks its bins is chosen so that the 32 lanes of a step
// read 32 different banks of the magnitude row.  Filter j is the sum of <= 4 partial sums (D of the cells of segment j, U of
// those of segment j - 1), combined through a 193-float scratch row per warp.
//   words: float4 [NP][32]  (wD_a, wU_a, wD_b, wU_b)      two steps per entry (a, b), NP = pairs[0] + pairs[1] + pairs[2]
//          u32    [NP][32]  byte offset of bin a | byte offset of bin b << 16
//          u32    [3][32]   four scratch indices (one byte each; 2 cell + {0: D, 1: U}; 192: the zero cell) of filter 32 r + lane
struct MelSegSchedule {
  bool ok = false;
  int pairs[3] = {0, 0, 0};
  std::vector<uint32_t> words;
};

inline MelSegSchedule mel_segment_schedule(const std::vector<double>& mel, int num_mels, int F) {
  MelSegSchedule out;
  if (num_mels < 1 || num_mels > 96 || F < 32 || F > 16000) return out;
  struct Tap { int bin; float wd, wu; };
  std::vector<std::vector<Tap>> seg(num_mels + 1);                 // seg[s + 1]: bins between the peaks of filters s and s + 1
  std::vector<int> peak(num_mels, 0);
  for (int m = 0; m < num_mels; ++m)
    for (int k = 1; k < F; ++k)
      if (mel[(size_t)m * F + k] > mel[(size_t)m * F + peak[m]]) peak[m] = k;
  for (int k = 0; k < F; ++k) {
    int f0 = -1, f1 = -1, cnt = 0;
    for (int m = 0; m < num_mels; ++m)
      if ((float)mel[(size_t)m * F + k] != 0.0f) { if (cnt == 0) f0 = m; else f1 = m; ++cnt; }
    if (cnt == 0) continue;
    if (cnt > 2 || (cnt == 2 && f1 != f0 + 1)) return out;         // not a 50 %-overlap bank: the lane schedule serves it
    const float w0 = (float)mel[(size_t)f0 * F + k];
    int s; Tap t{k, 0.f, 0.f};
    if (cnt == 2) { s = f0; t.wd = w0; t.wu = (float)mel[(size_t)f1 * F + k]; }
    else if (k > peak[f0]) { s = f0; t.wd = w0; }                  // only the falling side of f0 (the next filter starts later)
    else { s = f0 - 1; t.wu = w0; }                                // only its rising side (or its peak)
    seg[s + 1].push_back(t);
  }
  // split: the piece length that gives the fewest steps with <= 96 cells and <= 4 partial sums per filter
  size_t longest = 1;
  for (auto& sg : seg) longest = std::max(longest, sg.size());
  auto n_pieces = [&](size_t n, size_t C) { return (n + C - 1) / C; };
  size_t best_c = 0, best_steps = ~(size_t)0;
  for (size_t C = 2; C <= longest + 1; ++C) {
    std::vector<size_t> sizes;
    bool fits = true;
    for (int s = 0; s <= num_mels; ++s) {
      const size_t q = n_pieces(seg[s].size(), C);
      for (size_t i = 0; i < q; ++i) sizes.push_back(seg[s].size() / q + (i < seg[s].size() % q ? 1 : 0));
      if (s > 0 && q + n_pieces(seg[s - 1].size(), C) > 4) fits = false;
    }
    if (!fits || sizes.size() > 96) continue;
    std::sort(sizes.begin(), sizes.end(), [](size_t a, size_t b) { return a > b; });
    size_t steps = 0;
    for (size_t sl = 0; sl < 3; ++sl) if (sizes.size() > 32 * sl) steps += (sizes[32 * sl] + 1) / 2 * 2;
    if (steps < best_steps) { best_steps = steps; best_c = C; }
  }
  if (best_c == 0) return out;
  struct Cell { int s; std::vector<Tap> taps; };
  std::vector<Cell> cells;
  for (int s = 0; s <= num_mels; ++s) {
    const size_t n = seg[s].size(), q = n_pieces(n, best_c);
    size_t at = 0;
    for (size_t i = 0; i < q; ++i) {
      const size_t len = n / q + (i < n % q ? 1 : 0);
      cells.push_back({s - 1, std::vector<Tap>(seg[s].begin() + at, seg[s].begin() + at + len)});
      at += len;
    }
  }
  std::stable_sort(cells.begin(), cells.end(), [](const Cell& a, const Cell& b) { return a.taps.size() > b.taps.size(); });
  int first_pair[3] = {0, 0, 0}, np = 0;
  for (int sl = 0; sl < 3; ++sl) {
    first_pair[sl] = np;
    out.pairs[sl] = cells.size() > (size_t)32 * sl ? (int)((cells[32 * sl].taps.size() + 1) / 2) : 0;
    np += out.pairs[sl];
  }
  out.words.assign((size_t)160 * np + 96, 0u);
  auto put_f = [&](size_t i, float v) { uint32_t u; std::memcpy(&u, &v, 4); out.words[i] = u; };
  for (int sl = 0; sl < 3; ++sl) {
    const int steps = 2 * out.pairs[sl];
    std::vector<std::vector<char>> done(32);
    std::vector<int> left(32, 0);
    for (int l = 0; l < 32; ++l) {
      const size_t c = (size_t)32 * sl + l;
      if (c < cells.size()) { done[l].assign(cells[c].taps.size(), 0); left[l] = (int)cells[c].taps.size(); }
    }
    for (int st = 0; st < steps; ++st) {
      unsigned used = 0;                                            // banks read in this step
      std::vector<int> lanes(32);
      for (int l = 0; l < 32; ++l) lanes[l] = l;
      std::stable_sort(lanes.begin(), lanes.end(), [&](int x, int y) { return left[x] > left[y]; });   // lanes without slack choose first
      for (int l : lanes) {
        const size_t c = (size_t)32 * sl + l;
        int pick = -1;
        if (left[l] > 0) {
          const std::vector<Tap>& tp = cells[c].taps;
          for (size_t i = 0; i < tp.size(); ++i)
            if (!done[l][i] && !(used >> (tp[i].bin & 31) & 1u)) { pick = (int)i; break; }
          if (pick < 0 && left[l] >= steps - st)
            for (size_t i = 0; i < tp.size(); ++i) if (!done[l][i]) { pick = (int)i; break; }           // a bank conflict
        }
        int bin = 0; float wd = 0.f, wu = 0.f;
        if (pick >= 0) {
          const Tap& t = cells[c].taps[pick];
          done[l][pick] = 1; --left[l];
          bin = t.bin; wd = t.wd; wu = t.wu;
        } else {                                                    // idle step: zero weights, a bin in a free bank
          for (int b = 0; b < 32; ++b) if (!(used >> b & 1u)) { bin = b; break; }
        }
        used |= 1u << (bin & 31);
        const size_t pr = (size_t)first_pair[sl] + st / 2, h = st & 1;
        put_f((pr * 32 + l) * 4 + 2 * h, wd);
        put_f((pr * 32 + l) * 4 + 2 * h + 1, wu);
        out.words[(size_t)128 * np + pr * 32 + l] |= (uint32_t)(bin * 4) << (16 * h);
      }
    }
    for (int l = 0; l < 32; ++l) if (left[l] != 0) return out;
  }
  for (int j = 0; j < 96; ++j) {
    uint32_t w = 0; int n = 0;
    for (size_t c = 0; c < cells.size() && j < num_mels; ++c) {
      if (cells[c].s == j && n < 4) w |= (uint32_t)(2 * c) << (8 * n++);
      if (cells[c].s == j - 1 && n < 4) w |= (uint32_t)(2 * c + 1) << (8 * n++);
    }
    for (; n < 4; ++n) w |= 192u << (8 * n);
    out.words[(size_t)160 * np + j] = w;
  }
  out.ok = true;
  return out;
}

}  // namespace ttsa_host
