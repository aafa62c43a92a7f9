// Host-side (float64) construction of the plan's immutable tables: Hann window, twiddles, window-sum-square,
// Slaney mel basis (librosa.filters.mel semantics, utils/audio.py:68-77) and its Moore-Penrose pseudo-inverse
// (np.linalg.pinv, utils/audio.py:65).  Table construction only -- no signal data passes through here.
#pragma once
#include <algorithm>
#include <cmath>
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

}  // namespace ttsa_host
