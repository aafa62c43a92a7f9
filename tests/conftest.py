import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden", "ref_shim_golden.npz")

# audio blocks of the reference's config.json:5-25 and tests/test_config.json:2-21
MAIN_AUDIO = dict(num_mels=80, num_freq=1025, sample_rate=22050, frame_length_ms=50, frame_shift_ms=12.5,
                  preemphasis=0.98, min_level_db=-100, ref_level_db=20, power=1.5, griffin_lim_iters=60,
                  signal_norm=True, symmetric_norm=False, max_norm=1, clip_norm=True, mel_fmin=0.0,
                  mel_fmax=8000.0, do_trim_silence=True)
TEST_AUDIO = dict(audio_processor="audio", num_mels=80, num_freq=1025, sample_rate=22050, frame_length_ms=50,
                  frame_shift_ms=12.5, preemphasis=0.97, min_level_db=-100, ref_level_db=20, power=1.5,
                  griffin_lim_iters=30, signal_norm=True, symmetric_norm=True, clip_norm=True, max_norm=4,
                  mel_fmin=95, mel_fmax=7600, do_trim_silence=False)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden():
    g = np.load(GOLDEN)
    return {k: g[k] for k in g.files}


@pytest.fixture(scope="session")
def golden_audio_cfgs(golden):
    return {name: json.loads(bytes(golden[f"{name}_audio_json"]).decode()) for name in ("test", "main")}


def snr_db(ref, est):
    ref = np.asarray(ref, dtype=np.float64)
    est = np.asarray(est, dtype=np.float64)
    noise = np.sum((ref - est) ** 2)
    sig = np.sum(ref ** 2)
    if noise == 0:
        return np.inf
    return 10.0 * np.log10(sig / noise)


def synth_speech_like(seed, n_samples=132300, sr=22050):
    """Seeded speech-like test wave (harmonic stack with vibrato, AM envelope, noise floor)."""
    rng = np.random.default_rng(seed)
    t = np.arange(n_samples) / sr
    f0 = 120.0 + 30.0 * np.sin(2 * np.pi * 0.7 * t + rng.uniform(0, 2 * np.pi))
    phi = 2 * np.pi * np.cumsum(f0) / sr
    y = np.zeros(n_samples)
    for k in range(1, 31):
        y += np.sin(k * phi + rng.uniform(0, 2 * np.pi)) / k
    env = 0.55 + 0.45 * np.sin(2 * np.pi * 2.3 * t + rng.uniform(0, 2 * np.pi))
    y = 0.25 * y * env + 0.003 * rng.standard_normal(n_samples)
    return y.astype(np.float32)


def run_reference_test_normalize(ap, wav):
    """The assertion sequence of the reference's tests/test_audio.py:57-144, on any AudioProcessor-like ``ap``
    (attributes are mutated between calls exactly as the reference test does)."""
    ap.signal_norm = False
    x = np.asarray(ap.melspectrogram(wav), dtype=np.float64)
    x_old = x
    # (symmetric, clip, max_norm) in the order the reference walks them; None = leave clip_norm as it is
    for sym, clip, maxn in [(False, False, 4.0), (False, True, 4.0), (True, False, 4.0), (True, True, 4.0),
                            (False, None, 1.0), (True, None, 1.0)]:
        ap.signal_norm = True
        ap.symmetric_norm = sym
        if clip is not None:
            ap.clip_norm = clip
        ap.max_norm = maxn
        x_norm = np.asarray(ap._normalize(x), dtype=np.float64)
        assert (x_old - x).sum() == 0
        if ap.clip_norm:
            assert x_norm.max() <= ap.max_norm, x_norm.max()
            assert x_norm.min() >= (-ap.max_norm if sym else 0), x_norm.min()
        else:
            assert x_norm.max() <= ap.max_norm + 1, x_norm.max()
            assert x_norm.min() >= (-ap.max_norm - 2 if sym else 0 - 1), x_norm.min()
        if sym:
            assert x_norm.min() <= 0, x_norm.min()
        x_ = np.asarray(ap._denormalize(x_norm), dtype=np.float64)
        assert (x - x_).sum() < 1e-3, (x - x_).mean()
