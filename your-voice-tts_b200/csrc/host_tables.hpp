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

// Symmetric eigen-decomposition by cyclic Jacobi: A (n x n, overwritten) -> eigenvalues on the diagonal, V columns.
inline void jacobi_eigh(std::vector<double>& A, std::vector<double>& V, int n) {
  V.assign((size_t)n * n, 0.0);
  for (int i = 0; i < n; ++i) V[(size_t)i * n + i] = 1.0;
  for (int sweep = 0; sweep < 60; ++sweep) {
    double off = 0.0, diag = 0.0;
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n; ++j) (i == j ? diag : off) += A[(size_t)i * n + j] * A[(size_t)i * n + j];
    if (off <= 1e-30 * diag) break;
    for (int p = 0; p < n - 1; ++p) {
      for (int q = p + 1; q < n; ++q) {
        const double apq = A[(size_t)p * n + q];
        if (apq == 0.0) continue;
        const double app = A[(size_t)p * n + p], aqq = A[(size_t)q * n + q];
        const double theta = (aqq - app) / (2.0 * apq);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
        const double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < n; ++k) {   // columns p, q
          const double akp = A[(size_t)k * n + p], akq = A[(size_t)k * n + q];
          A[(size_t)k * n + p] = c * akp - s * akq;
          A[(size_t)k * n + q] = s * akp + c * akq;
        }
        for (int k = 0; k < n; ++k) {   // rows p, q
          const double apk = A[(size_t)p * n + k], aqk = A[(size_t)q * n + k];
          A[(size_t)p * n + k] = c * apk - s * aqk;
          A[(size_t)q * n + k] = s * apk + c * aqk;
        }
        for (int k = 0; k < n; ++k) {
          const double vkp = V[(size_t)k * n + p], vkq = V[(size_t)k * n + q];
          V[(size_t)k * n + p] = c * vkp - s * vkq;
          V[(size_t)k * n + q] = s * vkp + c * vkq;
        }
      }
    }
  }
}

// pinv(M) for M [m x n], m <= n:  M^+ = M^T (M M^T)^+ ; the Gram matrix is inverted through its eigen-decomposition
// with numpy's relative cut-off on the singular values (s_i > rcond * s_max, rcond = 1e-15).  Returns [n x m].
inline std::vector<double> pinv_wide(const std::vector<double>& M, int m, int n) {
  std::vector<double> G((size_t)m * m, 0.0), V;
  for (int i = 0; i < m; ++i)
    for (int j = i; j < m; ++j) {
      double acc = 0.0;
      for (int k = 0; k < n; ++k) acc += M[(size_t)i * n + k] * M[(size_t)j * n + k];
      G[(size_t)i * m + j] = G[(size_t)j * m + i] = acc;
    }
  jacobi_eigh(G, V, m);
  double lmax = 0.0;
  for (int i = 0; i < m; ++i) lmax = std::max(lmax, G[(size_t)i * m + i]);
  // eigenvalues of the Gram matrix are squared singular values; they are only resolved to ~1e-16 * lmax
  const double cut = std::max(1e-30, 1e-13) * lmax;
  std::vector<double> Ginv((size_t)m * m, 0.0);
  for (int e = 0; e < m; ++e) {
    const double lam = G[(size_t)e * m + e];
    if (lam <= cut) continue;
    for (int i = 0; i < m; ++i)
      for (int j = 0; j < m; ++j) Ginv[(size_t)i * m + j] += V[(size_t)i * m + e] * V[(size_t)j * m + e] / lam;
  }
  std::vector<double> P((size_t)n * m, 0.0);
  for (int k = 0; k < n; ++k)
    for (int j = 0; j < m; ++j) {
      double acc = 0.0;
      for (int i = 0; i < m; ++i) acc += M[(size_t)i * n + k] * Ginv[(size_t)i * m + j];
      P[(size_t)k * m + j] = acc;
    }
  return P;
}

}  // namespace ttsa_host
