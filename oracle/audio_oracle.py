"""CPU oracle for the spectrogram-domain audio hot path (TEST INFRASTRUCTURE ONLY).

This file is a float64 numpy restatement of the reference's
``utils/audio.py::AudioProcessor`` (reference file:line cited on every
function) together with the librosa-0.6.2 / scipy semantics that file calls
into.  librosa is a third-party dependency of the reference (pinned
``librosa==0.6.2`` in setup.py:76, ``==0.5.1`` in requirements.txt:3), it is
NOT vendored under /root/reference and cannot be installed here (no network),
so its published algorithm is restated below (functions prefixed ``lr_``).

PARITY PINNING STATUS
    The reference holds no golden vectors for this path (tests/test_audio.py
    asserts only normalisation ranges).  The oracle is therefore pinned by
      (1) the reference's own test_normalize assertions (tests/test_audio.py:57-144),
      (2) running the reference's *own* utils/audio.py code with its ``librosa``
          import satisfied by a torch.stft/torch.istft/torchaudio-backed shim
          (tests/golden/make_golden.py) -> committed fixtures under tests/golden/,
      (3) independent implementations of the same published algorithms:
          torch.stft / torch.istft (float64), torchaudio slaney filterbanks,
          scipy.signal.lfilter.
    At the librosa boundary itself parity is UNPINNED (librosa absent).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference
arm may import this module.  The product path never does.
"""
from __future__ import annotations

import numpy as np

__all__ = [
    "OracleAudioProcessor",
    "lr_hann_padded",
    "lr_stft",
    "lr_istft",
    "lr_window_sumsquare",
    "lr_mel",
    "reflect_index",
    "lfilter_fir2",
    "lfilter_iir1",
    "spectral_convergence",
]


# --------------------------------------------------------------------------
# librosa 0.6.2 semantics (restated; librosa is absent from /root/reference)
# --------------------------------------------------------------------------
def lr_hann_padded(win_length: int, n_fft: int) -> np.ndarray:
    """scipy.signal.get_window('hann', win, fftbins=True) centred in n_fft.

    librosa.stft / istft: ``fft_window = get_window(window, win_length,
    fftbins=True); util.pad_center(fft_window, n_fft)``  (left pad
    ``(n_fft - win)//2``).  Called from utils/audio.py:191-201.
    """
    n = np.arange(win_length, dtype=np.float64)
    w = 0.5 - 0.5 * np.cos(2.0 * np.pi * n / win_length)  # periodic Hann
    lpad = (n_fft - win_length) // 2
    out = np.zeros(n_fft, dtype=np.float64)
    out[lpad:lpad + win_length] = w
    return out


def reflect_index(i: np.ndarray, length: int) -> np.ndarray:
    """Index map of ``np.pad(y, pad, mode='reflect')`` (no edge repeat).

    Triangle-wave fold; equals numpy's iterated reflection for every
    length >= 2 (also when pad > length).  length == 1 maps to 0.
    """
    i = np.asarray(i, dtype=np.int64)
    if length == 1:
        return np.zeros_like(i)
    period = 2 * (length - 1)
    m = np.mod(i, period)
    return np.where(m >= length, period - m, m)


def lr_stft(y, n_fft: int, hop_length: int, win_length: int, dtype=np.complex128):
    """librosa.stft(y, n_fft, hop_length, win_length) with its defaults
    window='hann', center=True, pad_mode='reflect' (utils/audio.py:191-197).

    Returns [1 + n_fft//2, T] with T = 1 + (len(y) + 2*(n_fft//2) - n_fft)//hop.
    librosa stores complex64; ``dtype`` selects that storage truncation
    (strict oracle = complex128).
    """
    y = np.asarray(y, dtype=np.float64)
    w = lr_hann_padded(win_length, n_fft)
    yp = np.pad(y, n_fft // 2, mode="reflect")
    n_frames = 1 + (len(yp) - n_fft) // hop_length
    idx = np.arange(n_fft)[:, None] + hop_length * np.arange(n_frames)[None, :]
    frames = yp[idx] * w[:, None]
    return np.fft.rfft(frames, n=n_fft, axis=0).astype(dtype)


def lr_window_sumsquare(n_frames: int, hop_length: int, win_length: int, n_fft: int) -> np.ndarray:
    """librosa.filters.window_sumsquare('hann', n_frames, hop, win, n_fft, norm=None)."""
    n = n_fft + hop_length * (n_frames - 1)
    x = np.zeros(n, dtype=np.float64)
    wsq = lr_hann_padded(win_length, n_fft) ** 2
    for i in range(n_frames):
        s = i * hop_length
        x[s:min(n, s + n_fft)] += wsq[:max(0, min(n_fft, n - s))]
    return x


def lr_istft(stft_matrix, hop_length: int, win_length: int, dtype=np.float64):
    """librosa.istft(Y, hop_length, win_length) with defaults window='hann',
    center=True, length=None (utils/audio.py:199-201).

    Per frame: hermitian-extend, ifft().real (== irfft, Im of DC/Nyquist
    ignored), times padded window, overlap-add at t*hop; divide by
    window_sumsquare where > tiny; trim n_fft//2 at each end.  librosa
    accumulates in float32 (``dtype``); the strict oracle uses float64.
    """
    Y = np.asarray(stft_matrix)
    n_fft = 2 * (Y.shape[0] - 1)
    n_frames = Y.shape[1]
    w = lr_hann_padded(win_length, n_fft)
    expected = n_fft + hop_length * (n_frames - 1)
    y = np.zeros(expected, dtype=dtype)
    ytmp = w[:, None] * np.fft.irfft(Y, n=n_fft, axis=0)
    for t in range(n_frames):
        s = t * hop_length
        y[s:s + n_fft] += ytmp[:, t].astype(dtype)
    wss = lr_window_sumsquare(n_frames, hop_length, win_length, n_fft).astype(dtype)
    nz = wss > np.finfo(dtype).tiny
    y[nz] /= wss[nz]
    return y[n_fft // 2: expected - n_fft // 2]


def _hz_to_mel_slaney(f):
    f = np.asanyarray(f, dtype=np.float64)
    f_sp = 200.0 / 3
    mels = f / f_sp
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    safe = np.maximum(f, 1e-300)
    return np.where(f >= min_log_hz, min_log_mel + np.log(safe / min_log_hz) / logstep, mels)


def _mel_to_hz_slaney(m):
    m = np.asanyarray(m, dtype=np.float64)
    f_sp = 200.0 / 3
    freqs = f_sp * m
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), freqs)


def lr_mel(sr, n_fft: int, n_mels: int, fmin=0.0, fmax=None) -> np.ndarray:
    """librosa.filters.mel(sr, n_fft, n_mels, fmin, fmax, htk=False, norm=1)
    (called at utils/audio.py:72-77).  Slaney mel scale, area-normalised
    triangles, float64, shape [n_mels, 1 + n_fft//2]."""
    if fmax is None:
        fmax = float(sr) / 2
    n_mels = int(n_mels)
    n_bins = 1 + n_fft // 2
    fftfreqs = np.linspace(0, float(sr) / 2, n_bins, endpoint=True)
    lo, hi = _hz_to_mel_slaney(fmin), _hz_to_mel_slaney(fmax)
    mel_f = _mel_to_hz_slaney(np.linspace(lo, hi, n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = np.subtract.outer(mel_f, fftfreqs)
    weights = np.zeros((n_mels, n_bins), dtype=np.float64)
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        weights[i] = np.maximum(0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels])
    weights *= enorm[:, None]
    return weights


# --------------------------------------------------------------------------
# scipy.signal.lfilter restatements (utils/audio.py:128-136)
# --------------------------------------------------------------------------
def lfilter_fir2(x, p: float) -> np.ndarray:
    """signal.lfilter([1, -p], [1], x): y[n] = x[n] - p*x[n-1], x[-1] = 0."""
    x = np.asarray(x, dtype=np.float64)
    y = x.copy()
    y[1:] -= p * x[:-1]
    return y


def lfilter_iir1(x, p: float) -> np.ndarray:
    """signal.lfilter([1], [1, -p], x): y[n] = x[n] + p*y[n-1], y[-1] = 0."""
    x = np.asarray(x, dtype=np.float64)
    y = np.empty_like(x)
    acc = 0.0
    for n in range(len(x)):
        acc = x[n] + p * acc
        y[n] = acc
    return y


def spectral_convergence(mag_est, mag_target) -> float:
    """||  |X| - S ||_F / || S ||_F  (the per-iteration GL consistency measure)."""
    return float(np.linalg.norm(mag_est - mag_target) / max(np.linalg.norm(mag_target), 1e-300))


# --------------------------------------------------------------------------
# AudioProcessor restatement (utils/audio.py:11-201)
# --------------------------------------------------------------------------
class OracleAudioProcessor(object):
    """Line-by-line float64 restatement of utils/audio.py::AudioProcessor."""

    def __init__(self, sample_rate=None, num_mels=None, min_level_db=None, frame_shift_ms=None,
                 frame_length_ms=None, ref_level_db=None, num_freq=None, power=None, preemphasis=None,
                 signal_norm=None, symmetric_norm=None, max_norm=None, mel_fmin=None, mel_fmax=None,
                 clip_norm=True, griffin_lim_iters=None, do_trim_silence=False, **kwargs):
        # utils/audio.py:12-51
        self.sample_rate = sample_rate
        self.num_mels = num_mels
        self.min_level_db = min_level_db
        self.frame_shift_ms = frame_shift_ms
        self.frame_length_ms = frame_length_ms
        self.ref_level_db = ref_level_db
        self.num_freq = num_freq
        self.power = power
        self.preemphasis = preemphasis
        self.griffin_lim_iters = griffin_lim_iters
        self.signal_norm = signal_norm
        self.symmetric_norm = symmetric_norm
        self.mel_fmin = 0 if mel_fmin is None else mel_fmin
        self.mel_fmax = mel_fmax
        self.max_norm = 1.0 if max_norm is None else float(max_norm)
        self.clip_norm = clip_norm
        self.do_trim_silence = do_trim_silence
        self.n_fft, self.hop_length, self.win_length = self._stft_parameters()

    # utils/audio.py:114-119
    def _stft_parameters(self):
        n_fft = (self.num_freq - 1) * 2
        hop_length = int(self.frame_shift_ms / 1000.0 * self.sample_rate)
        win_length = int(self.frame_length_ms / 1000.0 * self.sample_rate)
        return n_fft, hop_length, win_length

    # utils/audio.py:68-77
    def _build_mel_basis(self):
        n_fft = (self.num_freq - 1) * 2
        if self.mel_fmax is not None:
            assert self.mel_fmax <= self.sample_rate // 2
        return lr_mel(self.sample_rate, n_fft, n_mels=self.num_mels, fmin=self.mel_fmin, fmax=self.mel_fmax)

    # utils/audio.py:60-62
    def _linear_to_mel(self, spectrogram):
        return np.dot(self._build_mel_basis(), spectrogram)

    # utils/audio.py:64-66
    def _mel_to_linear(self, mel_spec):
        inv_mel_basis = np.linalg.pinv(self._build_mel_basis())
        return np.maximum(1e-10, np.dot(inv_mel_basis, mel_spec))

    # utils/audio.py:79-94
    def _normalize(self, S):
        if self.signal_norm:
            S_norm = ((S - self.min_level_db) / - self.min_level_db)
            if self.symmetric_norm:
                S_norm = ((2 * self.max_norm) * S_norm) - self.max_norm
                if self.clip_norm:
                    S_norm = np.clip(S_norm, -self.max_norm, self.max_norm)
                return S_norm
            else:
                S_norm = self.max_norm * S_norm
                if self.clip_norm:
                    S_norm = np.clip(S_norm, 0, self.max_norm)
                return S_norm
        else:
            return S

    # utils/audio.py:96-112
    def _denormalize(self, S):
        S_denorm = S
        if self.signal_norm:
            if self.symmetric_norm:
                if self.clip_norm:
                    S_denorm = np.clip(S_denorm, -self.max_norm, self.max_norm)
                S_denorm = ((S_denorm + self.max_norm) * -self.min_level_db / (2 * self.max_norm)) + self.min_level_db
                return S_denorm
            else:
                if self.clip_norm:
                    S_denorm = np.clip(S_denorm, 0, self.max_norm)
                S_denorm = (S_denorm * -self.min_level_db / self.max_norm) + self.min_level_db
                return S_denorm
        else:
            return S

    # utils/audio.py:121-123
    def _amp_to_db(self, x):
        min_level = np.exp(self.min_level_db / 20 * np.log(10))
        return 20 * np.log10(np.maximum(min_level, x))

    # utils/audio.py:125-126
    def _db_to_amp(self, x):
        return np.power(10.0, x * 0.05)

    # utils/audio.py:128-131
    def apply_preemphasis(self, x):
        if self.preemphasis == 0:
            raise RuntimeError(" !! Preemphasis is applied with factor 0.0. ")
        return lfilter_fir2(x, self.preemphasis)

    # utils/audio.py:133-136
    def apply_inv_preemphasis(self, x):
        if self.preemphasis == 0:
            raise RuntimeError(" !! Preemphasis is applied with factor 0.0. ")
        return lfilter_iir1(x, self.preemphasis)

    # utils/audio.py:191-197
    def _stft(self, y):
        return lr_stft(y, self.n_fft, self.hop_length, self.win_length)

    # utils/audio.py:199-201
    def _istft(self, Y):
        return lr_istft(Y, self.hop_length, self.win_length)

    # utils/audio.py:138-144
    def spectrogram(self, y):
        y = np.asarray(y, dtype=np.float64)
        if self.preemphasis != 0:
            D = self._stft(self.apply_preemphasis(y))
        else:
            D = self._stft(y)
        S = self._amp_to_db(np.abs(D)) - self.ref_level_db
        return self._normalize(S)

    # utils/audio.py:146-152
    def melspectrogram(self, y):
        y = np.asarray(y, dtype=np.float64)
        if self.preemphasis != 0:
            D = self._stft(self.apply_preemphasis(y))
        else:
            D = self._stft(y)
        S = self._amp_to_db(self._linear_to_mel(np.abs(D))) - self.ref_level_db
        return self._normalize(S)

    # utils/audio.py:174-180
    def out_linear_to_mel(self, linear_spec):
        S = self._denormalize(np.asarray(linear_spec, dtype=np.float64))
        S = self._db_to_amp(S + self.ref_level_db)
        S = self._linear_to_mel(np.abs(S))
        S = self._amp_to_db(S) - self.ref_level_db
        return self._normalize(S)

    # utils/audio.py:182-189
    def _griffin_lim(self, S, init_angles=None, return_sc=False):
        """``init_angles`` [F, T] radians replaces ``2*pi*np.random.rand(*S.shape)``
        (utils/audio.py:183) so both sides of a parity test see the same phases.
        ``return_sc`` additionally returns, for iteration i = 1..iters,
        ||  |stft(y_{i-1})| - |S| ||_F / || S ||_F."""
        S = np.asarray(S, dtype=np.float64)
        if init_angles is None:
            init_angles = 2.0 * np.pi * np.random.rand(*S.shape)
        angles = np.exp(1j * np.asarray(init_angles, dtype=np.float64))
        S_complex = np.abs(S).astype(np.complex128)
        y = self._istft(S_complex * angles)
        sc = []
        for _ in range(self.griffin_lim_iters):
            X = self._stft(y)
            if return_sc:
                sc.append(spectral_convergence(np.abs(X), np.abs(S)))
            angles = np.exp(1j * np.angle(X))
            y = self._istft(S_complex * angles)
        if return_sc:
            return y, np.asarray(sc)
        return y

    # NOT in the reference: fast Griffin-Lim (Perraudin, Balazs, Sondergaard 2013) as librosa >= 0.7 implements it in
    # griffinlim(momentum=...); kept here as the checker of the product's opt-in momentum mode.
    def _griffin_lim_fast(self, S, momentum=0.99, init_angles=None, return_sc=False):
        """angles_i = angle(rebuilt_i - momentum / (1 + momentum) * rebuilt_{i-1}), rebuilt_i = stft(istft(S * angles_{i-1})),
        rebuilt_0 = 0; returns istft(S * angles_iters).  momentum = 0 is utils/audio.py:182-189 exactly.
        ``return_sc``: || |rebuilt_i - beta * rebuilt_{i-1}| - |S| ||_F / ||S||_F per iteration."""
        S = np.asarray(S, dtype=np.float64)
        if init_angles is None:
            init_angles = 2.0 * np.pi * np.random.rand(*S.shape)
        angles = np.exp(1j * np.asarray(init_angles, dtype=np.float64))
        S_complex = np.abs(S).astype(np.complex128)
        beta = momentum / (1.0 + momentum)
        rebuilt = 0.0
        sc = []
        for _ in range(self.griffin_lim_iters):
            tprev = rebuilt
            y = self._istft(S_complex * angles)
            rebuilt = self._stft(y)
            mixed = rebuilt - beta * tprev
            if return_sc:
                sc.append(spectral_convergence(np.abs(mixed), np.abs(S)))
            angles = np.exp(1j * np.angle(mixed))
        y = self._istft(S_complex * angles)
        if return_sc:
            return y, np.asarray(sc)
        return y

    # utils/audio.py:154-162
    def inv_spectrogram(self, spectrogram, init_angles=None, return_sc=False):
        S = self._denormalize(np.asarray(spectrogram, dtype=np.float64))
        S = self._db_to_amp(S + self.ref_level_db)
        out = self._griffin_lim(S ** self.power, init_angles, return_sc)
        y, sc = out if return_sc else (out, None)
        if self.preemphasis != 0:
            y = self.apply_inv_preemphasis(y)
        return (y, sc) if return_sc else y

    # utils/audio.py:164-172
    def inv_mel_spectrogram(self, mel_spectrogram, init_angles=None, return_sc=False):
        D = self._denormalize(np.asarray(mel_spectrogram, dtype=np.float64))
        S = self._db_to_amp(D + self.ref_level_db)
        S = self._mel_to_linear(S)
        out = self._griffin_lim(S ** self.power, init_angles, return_sc)
        y, sc = out if return_sc else (out, None)
        if self.preemphasis != 0:
            y = self.apply_inv_preemphasis(y)
        return (y, sc) if return_sc else y

    # utils/audio.py:56-58, up to the file write: the int16 samples save_wav hands to scipy.io.wavfile.write
    def save_wav_int16(self, wav):
        """float64 waveform (the reference's default path: scipy's lfilter in apply_inv_preemphasis and the server's
        list concatenation both widen): float64 arithmetic.  float32 waveform: under the reference's pinned numpy
        (1.14, value-based casting) the scale 32767 / max(0.01, peak) is a float64 scalar and the array product is
        formed in float32 with that scalar rounded to float32 -- written out explicitly so that the result does not
        depend on the numpy version running the oracle."""
        wav = np.asarray(wav)
        peak = float(np.max(np.abs(wav)))
        scale = 32767.0 / max(0.01, peak)
        if wav.dtype == np.float32:
            wav_norm = wav * np.float32(scale)
        else:
            wav_norm = wav.astype(np.float64) * scale
        return wav_norm.astype(np.int16)

    # server/synthesizer.py:157-161: every sentence is followed by 10 000 zero samples, then ONE save_wav
    @staticmethod
    def server_concat(wavs, gap=10000):
        out = []
        for w in wavs:
            out += list(w)
            out += [0] * gap
        return np.array(out)

    # utils/audio.py:203-210
    def find_endpoint(self, wav, threshold_db=-40, min_silence_sec=0.8):
        window_length = int(self.sample_rate * min_silence_sec)
        hop_length = int(window_length / 4)
        threshold = self._db_to_amp(threshold_db)
        for x in range(hop_length, len(wav) - window_length, hop_length):
            if np.max(wav[x:x + window_length]) < threshold:
                return x + hop_length
        return len(wav)

    # utils/audio.py:220-235
    @staticmethod
    def mulaw_encode(wav, qc):
        mu = 2 ** qc - 1
        signal = np.sign(wav) * np.log(1 + mu * np.abs(wav)) / np.log(1. + mu)
        signal = (signal + 1) / 2 * mu + 0.5
        return np.floor(signal)

    @staticmethod
    def mulaw_decode(wav, qc):
        mu = 2 ** qc - 1
        return np.sign(wav) / mu * ((1 + mu) ** np.abs(wav) - 1)
